"""Seeded random stress of the certification logic: stores made of near-duplicate clusters whose spread is drawn
around the scan's error bound (the hard regime: bands that hold a few, dozens or hundreds of rows), random shapes,
dtypes, batch sizes and k.  Every result must equal the oracle's -- index lists and binary64 scores -- and the scan's
error bound must hold on every rescored candidate."""
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


def _store_values(X, dtype):
    """the values the store scores: rounded once for fp32 / bf16 stores, as given for binary64 stores"""
    import torch
    if dtype.startswith("f64"):
        return X
    if dtype == "bf16":
        return torch.from_numpy(X).to(torch.bfloat16).to(torch.float64).numpy()
    return X.astype(np.float32).astype(np.float64)


CASES = list(range(48))


@pytest.mark.parametrize("case", CASES)
def test_random_clustered_store(vm, case):
    rng = np.random.default_rng(1000 + case)
    dtype = ["f32", "bf16", "f64", "f64+bf16"][case % 4]
    d = int(rng.choice([64, 128, 384, 768]))
    n = int(rng.integers(9_500, 140_000))
    per = int(rng.choice([1, 4, 16, 64, 200]))                        # rows per cluster (1: no clusters)
    spread = float(10.0 ** rng.uniform(-5.0, -0.5))                   # relative noise inside a cluster
    nq = int(rng.integers(1, 65))
    k = int(rng.choice([1, 3, 10, 16, 24, 40]))
    centres = rng.standard_normal((n // per + 1, d))
    X = centres[np.arange(n) // per] * (1.0 + rng.random((n, 1))) + spread * rng.standard_normal((n, d))
    if per > 1:
        X[n // 2: n // 2 + min(per, 50)] = X[n // 2]                  # exact duplicates inside one cluster
    X[7] = 0.0
    pick = rng.integers(0, n, nq)
    Q = X[pick] + spread * rng.standard_normal((nq, d))               # queries sit inside clusters
    Q[nq // 2] = rng.standard_normal(d)
    st = vm.EmbeddingStore(d, n, dtype)
    st.append(X)
    Xs = _store_values(X, dtype)
    plain = dtype.startswith("f64") or n * nq <= 2_000_000
    naive = plain and case % 3 == 0                                  # CPython < 3.12 summation order on a third of the cases
    ok = np.ones(n, np.uint8)
    if plain and case % 2 == 1:                                       # skipped rows, among them members of the queried clusters
        bad = np.unique(np.concatenate([rng.integers(0, n, 40), pick[: max(1, nq // 3)]]))
        ok[bad] = 0
        st.invalidate(bad)
    if plain:
        ref = oracle.batch_similarities(Q, Xs, k, row_ok=ok, sum_mode=oracle.SUM_NAIVE if naive else oracle.SUM_NEUMAIER)
        ri = np.array([[r for r, _ in lst] + [-1] * (k - len(lst)) for lst in ref])
        rs = np.array([[s for _, s in lst] + [0.0] * (k - len(lst)) for lst in ref])
    else:                                                             # blocked tier: float64 BLAS pre-ranking + bit-exact rescoring
        Q = Q.astype(np.float32).astype(np.float64)                   # (that tier takes float32 queries)
        ri, rs, _ = oracle.topk_blocked(Q.astype(np.float32), Xs.astype(np.float32), k, slack=max(24, 3 * per))
    idx, score, count = st.topk(Q, k, sum_mode=vm.VM_SUM_NAIVE if naive else vm.VM_SUM_NEUMAIER)
    assert (count == k).all()
    assert np.array_equal(idx, ri), (case, dtype, d, n, per, spread, nq, k)
    assert np.array_equal(score, rs), (case, dtype, d, n, per, spread, nq, k)
    c = st.counters()
    assert c["bound_violations"] == 0, c
    st.close()
