"""Parity of the SHARDED top-k path on a single GPU (the driver's GPU-test box has one device, so the 2-GPU
NCCL test is skipped there): G shard stores are built on the one device, each shard runs the same local pass
vm_topk_sharded runs on a rank (exact local top-k with its row_offset, results written straight into that rank's
block of the packed exchange buffer), and the blocks -- laid out exactly as the ncclAllGather receive buffer,
vm_topk_packed_bytes(nq, k) per rank -- go through the same merge kernel (vm_merge_topk_packed).  Checked against
the oracle over the UNSHARDED store: identical rows, bit-identical binary64 scores, ties -> lowest global row
also when the tied rows sit on different shards."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def vm():
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    import vidmem_b200
    vidmem_b200._lib.load()
    return vidmem_b200


@pytest.fixture(scope="module")
def comm1(vm):
    from vidmem_b200.sharded import Communicator
    c = Communicator(0, 1, 0, Communicator.unique_id())      # a one-rank NCCL communicator: row_offset is honoured
    yield c
    c.close()


def _sharded_topk(vm, comm1, X_store, Q, k, G, dtype, flags=0, min_score=-np.inf, score_mode=0):
    import torch
    from vidmem_b200.sharded import shard_bounds
    lib = vm._lib.load()
    n, d = X_store.shape
    nq = len(Q)
    per = int(lib.vm_topk_packed_bytes(nq, k))
    seg = nq * k * 8
    assert per == 2 * seg + ((nq * 4 + 7) & ~7)
    packed = torch.zeros((G * per,), dtype=torch.uint8, device="cuda")
    qd = torch.from_numpy(Q).cuda()
    stores = []
    for g, (lo, hi) in enumerate(shard_bounds(n, G)):
        st = vm.EmbeddingStore(d, max(hi - lo, 1), dtype)
        if hi > lo:
            st.append(X_store[lo:hi])
        blk = packed[g * per:(g + 1) * per]
        out = (blk[:seg].view(torch.int64).view(nq, k), blk[seg:2 * seg].view(torch.float64).view(nq, k),
               blk[2 * seg:2 * seg + nq * 4].view(torch.int32))
        st.topk_device(qd, k, out=out, comm=comm1, row_offset=lo, sum_mode=vm.VM_SUM_NEUMAIER, flags=flags | vm.VM_FLAG_ASYNC,
                       min_score=min_score, score_mode=score_mode)
        stores.append(st)
    oi = torch.empty((nq, k), dtype=torch.int64, device="cuda")
    os_ = torch.empty((nq, k), dtype=torch.float64, device="cuda")
    oc = torch.empty((nq,), dtype=torch.int32, device="cuda")
    vm._lib.check(lib.vm_merge_topk_packed(0, packed.data_ptr(), G, nq, k, oi.data_ptr(), os_.data_ptr(), oc.data_ptr(),
                                           torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    for st in stores:
        st.close()
    return oi.cpu().numpy(), os_.cpu().numpy(), oc.cpu().numpy()


def _quantise(x, dtype):
    if dtype == "bf16":
        import torch
        return torch.from_numpy(x).to(torch.bfloat16).to(torch.float32).numpy()
    return x


@pytest.mark.parametrize("G", [2, 4, 8])
@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_shards_on_one_gpu_match_the_unsharded_oracle(vm, comm1, G, dtype):
    n, d, nq, k = 30011, 384, 64, 10
    rng = np.random.default_rng(100 + G)
    X = _quantise(rng.standard_normal((n, d)).astype(np.float32), dtype)      # non-representable values, not multiples of 1/128
    X[n - 1] = X[2]                                                           # a tie between the first and the last shard
    X[n // 2 + 5] = X[2]                                                      # ... and a middle one
    X[7] = 0.0                                                                # a zero row scores 0.0
    Q = rng.standard_normal((nq, d)).astype(np.float32)
    Q[0] = X[2]
    Q[1] = X[n - 3] * 0.5
    Q[2] = 0.0                                                                # zero query: every score 0.0 -> first k global rows
    idx, score, count = _sharded_topk(vm, comm1, X, Q, k, G, dtype)
    ref = oracle.batch_similarities(Q, X, k)
    for qi, lst in enumerate(ref):
        assert count[qi] == len(lst)
        assert list(idx[qi, :len(lst)]) == [r for r, _ in lst], (qi, idx[qi], lst)
        assert list(score[qi, :len(lst)]) == [s for _, s in lst], qi
    assert list(idx[0, :3]) == [2, n // 2 + 5, n - 1]                         # equal scores across shards: lowest global row first
    assert list(idx[2]) == list(range(k))


def test_shards_with_threshold_short_lists_and_empty_shard(vm, comm1):
    """Fewer than k hits per shard (strict min_score on the Neo4j-normalised score), a shard with no rows at all
    (more ranks than rows), and k larger than a shard."""
    n, d, nq, k = 13, 64, 5, 6
    X = synth.synth_rows(71, 0, n, d)
    Q = synth.synth_queries(72, nq, d, 71, n)
    for G in (2, 8, 16):
        idx, score, count = _sharded_topk(vm, comm1, X, Q, k, G, "f32")
        ref = oracle.batch_similarities(Q, X, k)
        for qi, lst in enumerate(ref):
            assert count[qi] == len(lst) and list(idx[qi, :len(lst)]) == [r for r, _ in lst]
            assert list(score[qi, :len(lst)]) == [s for _, s in lst]
        idx, score, count = _sharded_topk(vm, comm1, X, Q, k, G, "f32", min_score=0.55, score_mode=vm.VM_SCORE_NEO4J)
        for qi in range(nq):
            want = oracle.vector_search(Q[qi], X.astype(np.float64), k, min_score=0.55)
            assert count[qi] == len(want) and list(idx[qi, :len(want)]) == [r for r, _ in want]
            assert list(score[qi, :len(want)]) == [s for _, s in want]
