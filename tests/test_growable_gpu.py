"""Growable stores: virtual addresses reserved once, physical HBM mapped behind the resident rows as
they arrive -- base addresses never move, nothing is copied, answers equal those of a store over fixed buffers and the
oracle's."""
import numpy as np
import pytest

from oracle import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def vm():
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    import vidmem_b200
    vidmem_b200._lib.load()
    return vidmem_b200


def _check(idx, score, count, ref):
    for qi, lst in enumerate(ref):
        assert count[qi] == len(lst)
        assert list(idx[qi, :len(lst)]) == [r for r, _ in lst], (qi, idx[qi], lst)
        assert list(score[qi, :len(lst)]) == [s for _, s in lst], qi


@pytest.mark.parametrize("dtype", ["f32", "bf16", "f64"])
def test_growth_in_place(vm, dtype):
    import torch
    d, n, k = 384, 300_000, 10
    rng = np.random.default_rng(41)
    X = rng.standard_normal((n, d)).astype(np.float32)
    Q = rng.standard_normal((7, d)).astype(np.float32)
    Q[0] = X[299_999] + 0.1 * rng.standard_normal(d).astype(np.float32)
    st = vm.EmbeddingStore(d, 1000, dtype, max_capacity=2_000_000)
    assert st.growable and st.capacity == 1000 and st.max_capacity == 2_000_000
    base = st.rows.data_ptr()
    small = st.resident_bytes()
    fixed = vm.EmbeddingStore(d, n, dtype)                        # the same rows in a store over fixed buffers
    torch.cuda.synchronize()
    free0 = torch.cuda.mem_get_info()[0]
    pos, step = 0, 777
    while pos < n:                                                 # ever larger appends: 777, 1554, 3108, ...
        m = min(step, n - pos)
        assert st.append(X[pos:pos + m]) == pos
        pos += m; step *= 2
        assert st.rows.data_ptr() == base and len(st) == pos and st.capacity >= pos
    fixed.append(X)
    row_bytes = st.ld * {"f32": 4, "bf16": 2, "f64": 12}[dtype] + 4
    assert small < st.resident_bytes() <= 2 * n * row_bytes + (256 << 20)      # backed: at most the doubling, never max_capacity
    # device memory actually taken: the backed rows + the library's upload staging buffer (largest append), no second copy
    assert free0 - torch.cuda.mem_get_info()[0] <= st.resident_bytes() + n * d * 4 + (96 << 20)
    a = st.topk(Q, k, sum_mode=vm.VM_SUM_NEUMAIER)
    b = fixed.topk(Q, k, sum_mode=vm.VM_SUM_NEUMAIER)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    assert a[0][0, 0] == 299_999
    if dtype != "bf16":
        Xo = X if dtype == "f32" else X.astype(np.float64)
        _check(*st.topk(Q[:3], k, sum_mode=vm.VM_SUM_NEUMAIER), oracle.batch_similarities(Q[:3], Xo, k))
    # rows written through the torch view are the store's rows (same memory)
    assert torch.equal(st.rows[:1000].float().cpu(), fixed.rows[:1000].float().cpu())
    # explicit reserve, then the limit
    st.reserve(500_000)
    assert st.capacity == 500_000 and st.rows.data_ptr() == base and st.rows.shape[0] == 500_000
    with pytest.raises(vm._lib.VidmemError):
        st.reserve(2_000_001)
    with pytest.raises(ValueError):
        fixed.reserve(n + 1)
    st.close(); fixed.close()


def test_resident_chunk_store_grows_without_copy(vm):
    """The adapter's store starts at its initial capacity and is fed batch after batch (the insert hook S6); the row
    buffer keeps its address through every growth step and the answers equal the oracle's on everything inserted."""
    from vidmem_b200 import adapters
    d, k = 64, 5
    rng = np.random.default_rng(6)
    rs = adapters.ResidentChunkStore("f32", 0, initial_capacity=256)
    all_rows, base = [], None
    for b in range(12):
        m = 100 * (b + 1)
        E = rng.standard_normal((m, d)).astype(np.float32)
        rs.upsert([(f"c{b}_{i}", E[i].tolist()) for i in range(m)])
        all_rows.append(E)
        if base is None:
            base = rs.store.rows.data_ptr()
        assert rs.store.rows.data_ptr() == base
    X = np.concatenate(all_rows)
    assert len(rs) == len(X) == 7800 and rs.store.capacity >= 7800 and rs.store.growable
    Q = rng.standard_normal((4, d)).astype(np.float32)
    _check(*rs.store.topk(Q, k, sum_mode=vm.VM_SUM_NEUMAIER), oracle.batch_similarities(Q, X, k))
    rs.store.close()
