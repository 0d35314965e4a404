"""GPU parity tests of the top-k scorer, through the C ABI (libvidmem.so), against the CPU
oracle and the committed golden fixtures.  Bar: index lists identical, binary64 scores
bit-identical (the engine rescoring runs the reference's own summation order)."""
import math
import os

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


def _check(idx, score, count, ref, k):
    for qi, lst in enumerate(ref):
        assert count[qi] == len(lst), (qi, count[qi], len(lst))
        assert list(idx[qi, :len(lst)]) == [r for r, _ in lst], (qi, idx[qi], lst)
        assert list(score[qi, :len(lst)]) == [s for _, s in lst], qi  # bit-exact binary64
        assert (idx[qi, len(lst):] == -1).all()


def _quantise(x, dtype):
    if dtype == "bf16":
        import torch
        return torch.from_numpy(x).to(torch.bfloat16).to(torch.float32).numpy()
    return x


@pytest.mark.parametrize("flags_name", ["default", "exact", "simt"])
@pytest.mark.parametrize("k", [3, 10])
def test_golden_batch_small(vm, golden_dir, k, flags_name):
    g = np.load(os.path.join(golden_dir, "batch_small.npz"))
    X, Q = g["X"], g["Q"]
    flags = {"default": 0, "exact": vm.VM_FLAG_FORCE_EXACT, "simt": vm.VM_FLAG_FORCE_SIMT}[flags_name]
    st = vm.EmbeddingStore(X.shape[1], 1024, "f32")
    st.append(X.astype(np.float64))
    st.invalidate(np.nonzero(g["row_ok"] == 0)[0])
    ok = np.nonzero(g["query_ok"])[0]          # Exception-valued queries never reach the engine (adapter -> [])
    idx, score, count = st.topk(Q[ok].astype(np.float64), k, sum_mode=vm.VM_SUM_NEUMAIER, flags=flags)
    for out_i, qi in enumerate(ok):
        c = int(g[f"count_k{k}"][qi])
        assert count[out_i] == c
        assert np.array_equal(idx[out_i, :c], g[f"idx_k{k}"][qi, :c])
        assert np.array_equal(score[out_i, :c], g[f"score_k{k}"][qi, :c])
    st.close()


def test_golden_c1(vm, golden_dir):
    g = np.load(os.path.join(golden_dir, "batch_c1.npz"))
    n, d, q, k = int(g["n"]), int(g["d"]), int(g["q"]), int(g["k"])
    X = synth.synth_rows(int(g["store_seed"]), 0, n, d)
    Q = synth.synth_queries(int(g["query_seed"]), q, d, int(g["store_seed"]), n)
    st = vm.EmbeddingStore(d, n, "f32")
    st.append(X)
    for flags in (0, vm.VM_FLAG_FORCE_SIMT, vm.VM_FLAG_FORCE_EXACT):
        idx, score, count = st.topk(Q, k, sum_mode=vm.VM_SUM_NEUMAIER, flags=flags)
        assert (count == k).all()
        assert np.array_equal(idx, g["idx"]), flags
        assert np.array_equal(score, g["score"]), flags
    # device-side generator must reproduce the host generator bit for bit
    st2 = vm.EmbeddingStore(d, n, "f32")
    st2.synth_fill(int(g["store_seed"]), n)
    st2.set_size(n)
    idx2, score2, _ = st2.topk(Q, k, sum_mode=vm.VM_SUM_NEUMAIER)
    assert np.array_equal(idx2, g["idx"]) and np.array_equal(score2, g["score"])
    st.close(); st2.close()


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
@pytest.mark.parametrize("n,d,nq,k", [(1, 8, 1, 1), (7, 24, 3, 10), (1000, 384, 5, 10), (4097, 768, 8, 3),
                                      (20000, 100, 2, 24), (3000, 384, 17, 10)])
def test_simt_vs_oracle(vm, dtype, n, d, nq, k):
    rng = np.random.default_rng(n * 31 + d)
    X = _quantise(rng.standard_normal((n, d)).astype(np.float32), dtype)
    Q = rng.standard_normal((nq, d)).astype(np.float32)
    if n > 10:
        X[3] = X[1]; X[n - 1] = X[1]; X[5] = 0.0
        Q[0] = X[1]
    st = vm.EmbeddingStore(d, n + 5, dtype)
    st.append(X)
    for sm, osm in ((vm.VM_SUM_NEUMAIER, oracle.SUM_NEUMAIER), (vm.VM_SUM_NAIVE, oracle.SUM_NAIVE)):
        ref = oracle.batch_similarities(Q, X, k, sum_mode=osm)
        for flags in (vm.VM_FLAG_FORCE_SIMT, vm.VM_FLAG_FORCE_EXACT):
            if flags == vm.VM_FLAG_FORCE_SIMT and k > 24:
                continue
            idx, score, count = st.topk(Q, k, sum_mode=sm, flags=flags)
            _check(idx, score, count, ref, k)
    st.close()


def test_many_duplicates_fall_back_to_exact(vm):
    # more exact ties at the top than the candidate list holds -> the whole tie band is rescored (300 duplicates) or,
    # the zero query (every row ties at 0.0) is answered from store order directly
    d, n, k = 64, 5000, 10
    rng = np.random.default_rng(5)
    X = rng.standard_normal((n, d)).astype(np.float32)
    X[100:400] = X[100]
    Q = np.stack([X[100], rng.standard_normal(d).astype(np.float32), np.zeros(d, np.float32)])
    st = vm.EmbeddingStore(d, n, "f32")
    st.append(X)
    idx, score, count = st.topk(Q, k, sum_mode=vm.VM_SUM_NEUMAIER)
    c = st.counters()
    assert c["uncertified"] >= 1 and c["band_settled"] >= 1 and c["full_rescans"] == 0 and c["bound_violations"] == 0
    _check(idx, score, count, oracle.batch_similarities(Q, X, k), k)
    assert list(idx[0]) == list(range(100, 110))
    assert list(idx[2]) == list(range(10)) and (score[2] == 0.0).all()   # zero query: store order
    st.close()


def test_min_score_and_neo4j_mode(vm):
    d, n = 32, 300
    X = synth.synth_rows(9, 0, n, d)
    q = X[7:8].copy()
    st = vm.EmbeddingStore(d, n, "f32")
    st.append(X)
    idx, score, count = st.topk(q, 5, min_score=0.3, score_mode=vm.VM_SCORE_NEO4J, sum_mode=vm.VM_SUM_NEUMAIER)
    ref = oracle.vector_search(q[0], X, 5, 0.3)
    assert count[0] == len(ref) and [(int(i), float(s)) for i, s in zip(idx[0, :count[0]], score[0, :count[0]])] == ref
    idx, score, count = st.topk(q, 5, min_score=0.999, sum_mode=vm.VM_SUM_NEUMAIER)
    assert count[0] == 1 and idx[0, 0] == 7 and idx[0, 1] == -1
    st.close()


def test_update_invalidate_append_visibility(vm):
    d = 48
    X = synth.synth_rows(21, 0, 200, d)
    st = vm.EmbeddingStore(d, 400, "f32")
    assert st.append(X[:100]) == 0
    assert st.append(X[100:]) == 100 and len(st) == 200
    q = X[150:151]
    idx, score, _ = st.topk(q, 1, sum_mode=vm.VM_SUM_NEUMAIER)
    assert idx[0, 0] == 150 and score[0, 0] == oracle.cosine(q[0], X[150])
    st.update(150, X[0:1])                  # upsert overwrites the row
    X2 = X.copy(); X2[150] = X[0]
    _check(*st.topk(q, 3, sum_mode=vm.VM_SUM_NEUMAIER), oracle.batch_similarities(q, X2, 3), 3)
    st.invalidate([int(idx[0, 0])])
    ok = np.ones(200, np.uint8); ok[150] = 0
    _check(*st.topk(q, 3, sum_mode=vm.VM_SUM_NEUMAIER), oracle.batch_similarities(q, X2, 3, row_ok=ok), 3)
    with pytest.raises(vm.VidmemError):
        st.append(np.zeros((500, d), np.float32))   # capacity overflow is loud
    st.close()


def test_device_api_async(vm):
    import torch
    d, n, nq, k = 384, 6000, 6, 10
    X = synth.synth_rows(31, 0, n, d)
    Q = synth.synth_queries(32, nq, d, 31, n)
    st = vm.EmbeddingStore(d, n, "bf16")
    st.append(torch.from_numpy(X).cuda())
    qd = torch.from_numpy(Q).cuda()
    idx, score, count = st.topk_device(qd, k, sum_mode=vm.VM_SUM_NEUMAIER, flags=vm.VM_FLAG_ASYNC)
    torch.cuda.synchronize()
    _check(idx.cpu().numpy(), score.cpu().numpy(), count.cpu().numpy(), oracle.batch_similarities(Q, X, k), k)
    st.close()


def test_cosine_pairs_golden(vm, golden_dir):
    from vidmem_b200.store import cosine_pairs
    g = np.load(os.path.join(golden_dir, "cosine_kat.npz"))
    for i in range(len(g["out"])):
        la, lb = int(g["len_a"][i]), int(g["len_b"][i])
        if la != lb:
            continue  # length mismatch is decided in the adapter, not on the device
        a, b = g["a"][i, :la], g["b"][i, :lb]
        assert cosine_pairs(a, b, zero_rule=0, sum_mode=vm.VM_SUM_NEUMAIER)[0] == g["out"][i, 0]
        assert cosine_pairs(a, b, zero_rule=1, sum_mode=vm.VM_SUM_NEUMAIER)[0] == g["out"][i, 1]
        # a4, EmbeddingUtils.cosine_similarity: `** 0.5` is libm pow in CPython and CUDA's pow here -- the one seam
        # with a tolerance (a few ulp of binary64; north_star allows 1e-5 relative)
        got = cosine_pairs(a, b, zero_rule=2, sum_mode=vm.VM_SUM_NEUMAIER)[0]
        assert got == pytest.approx(g["out"][i, 2], rel=4 * np.finfo(np.float64).eps, abs=0.0), i
        assert (got == 0.0) == (g["out"][i, 2] == 0.0)


def test_cosine_pairs_scratch_is_per_device(vm):
    """vm_cosine_pairs / vm_merge_topk_lists keep their device scratch per device (round-1 ADVICE: a buffer allocated
    on device A must not be handed to a kernel on device B); with one GPU this checks the per-device table and the
    rejection of an out-of-range index."""
    from vidmem_b200.store import cosine_pairs
    a, b = np.arange(1.0, 9.0), np.arange(8.0, 0.0, -1.0)
    want = oracle.cosine(a, b, "injector")
    for _ in range(3):
        assert cosine_pairs(a, b, device=0)[0] == want
    import torch
    if torch.cuda.device_count() > 1:
        assert cosine_pairs(a, b, device=1)[0] == want
        assert cosine_pairs(a, b, device=0)[0] == want
    with pytest.raises(Exception):
        cosine_pairs(a, b, device=99)


def test_merge_max_by_id_golden(vm, golden_dir):
    import ctypes as C
    import torch
    g = np.load(os.path.join(golden_dir, "merge.npz"))
    X, Q = g["X"], g["Q"]
    st = vm.EmbeddingStore(X.shape[1], len(X), "f32")
    st.append(X)
    lib = vm._lib.load()
    for k, k2 in ((3, 2), (10, 4), (10, 25)):
        idx, score, count = st.topk_device(torch.from_numpy(Q).cuda(), k, sum_mode=vm.VM_SUM_NEUMAIER)
        oi = torch.full((k2,), -1, dtype=torch.int64, device="cuda")
        os_ = torch.zeros((k2,), dtype=torch.float64, device="cuda")
        oc = torch.zeros((1,), dtype=torch.int32, device="cuda")
        vm._lib.check(lib.vm_merge_max_by_id(0, idx.data_ptr(), score.data_ptr(), count.data_ptr(), len(Q), k, k2,
                                             oi.data_ptr(), os_.data_ptr(), oc.data_ptr(),
                                             torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        m = int(oc.item())
        assert m == len(g[f"idx_{k}_{k2}"])
        assert np.array_equal(oi.cpu().numpy()[:m], g[f"idx_{k}_{k2}"])
        assert np.array_equal(os_.cpu().numpy()[:m], g[f"score_{k}_{k2}"])
    st.close()


# ---- tcgen05 / TMA scan kernel ----------------------------------------------------------------
@pytest.mark.parametrize("dtype", ["f32", "bf16"])
@pytest.mark.parametrize("n,d,nq,k", [(1, 8, 1, 1), (100, 384, 5, 10), (129, 64, 16, 3), (5000, 384, 64, 10),
                                      (4097, 768, 17, 10), (20000, 100, 33, 16), (60000, 384, 64, 10),
                                      (3000, 1024, 8, 40)])
def test_tc_vs_oracle(vm, dtype, n, d, nq, k):
    lib = vm._lib.load()
    rng = np.random.default_rng(n * 17 + d + nq)
    X = _quantise(rng.standard_normal((n, d)).astype(np.float32), dtype)
    Q = rng.standard_normal((nq, d)).astype(np.float32)
    if n > 10:
        X[3] = X[1]; X[n - 1] = X[1]; X[5] = 0.0
        Q[0] = X[1]
    st = vm.EmbeddingStore(d, n + 5, dtype)
    st.append(X)
    ref = oracle.batch_similarities(Q, X, k)
    idx, score, count = st.topk(Q, k, sum_mode=vm.VM_SUM_NEUMAIER, flags=vm.VM_FLAG_FORCE_TC)
    assert st.last_stats.scan_kernel == 2
    if n >= 3000:
        assert st.last_stats.uncertified <= 1  # only the planted-duplicate query may need the exact re-scan
    _check(idx, score, count, ref, k)
    st.close()


def test_tc_is_used_for_query_batches(vm):
    d, n, nq, k = 384, 40000, 64, 10
    X = synth.synth_rows(61, 0, n, d)
    Q = synth.synth_queries(62, nq, d, 61, n)
    for dtype in ("f32", "bf16"):
        st = vm.EmbeddingStore(d, n, dtype)
        st.append(X)
        idx, score, count = st.topk(Q, k, sum_mode=vm.VM_SUM_NEUMAIER)
        assert st.last_stats.scan_kernel == 2 and st.last_stats.scan_stages >= 3
        # the tensor-core candidates themselves must be right: nothing may lean on the exact re-scan
        assert st.last_stats.uncertified == 0
        _check(idx, score, count, oracle.batch_similarities(Q, X, k), k)
        # skipped rows and rows beyond the shard never surface from the tensor-core path
        st.invalidate([int(idx[0, 0]), int(idx[1, 0])])
        ok = np.ones(n, np.uint8); ok[[int(idx[0, 0]), int(idx[1, 0])]] = 0
        _check(*st.topk(Q, k, sum_mode=vm.VM_SUM_NEUMAIER), oracle.batch_similarities(Q, X, k, row_ok=ok), k)
        st.close()


@pytest.mark.parametrize("dtype,d,nq", [("f32", 1536, 64), ("f32", 768, 40), ("bf16", 1536, 64), ("f32", 1024, 70)])
def test_large_dims_split_query_batches(vm, dtype, d, nq):
    """64 queries of a 768..1536-d fp32 store do not fit the tensor-core kernel's shared memory at once:
    the engine splits the batch instead of leaving the tcgen05 path."""
    n, k = 70000, 10
    rng = np.random.default_rng(d + nq)
    X = _quantise(rng.standard_normal((n, d)).astype(np.float32), dtype)
    Q = rng.standard_normal((nq, d)).astype(np.float32)
    Q[0] = X[4242]
    st = vm.EmbeddingStore(d, n, dtype)
    st.append(X)
    idx, score, count = st.topk(Q, k, sum_mode=vm.VM_SUM_NEUMAIER)
    assert st.last_stats.scan_kernel == 2
    bi, bs, bc = oracle.topk_blocked(Q, X, k)
    assert np.array_equal(idx, bi) and np.array_equal(score, bs) and (count == k).all()
    assert idx[0, 0] == 4242
    st.close()


def test_save_load_sidecar_roundtrip(vm, tmp_path):
    d, n = 96, 500
    X = synth.synth_rows(55, 0, n, d)
    Q = synth.synth_queries(56, 4, d, 55, n)
    for dtype in ("f32", "bf16"):
        st = vm.EmbeddingStore(d, n, dtype)
        st.append(X)
        st.invalidate([3, 77])
        ref = st.topk(Q, 5)
        p = str(tmp_path / f"store_{dtype}")
        st.save(p, ids=[f"c{i}" for i in range(n)])
        st2, ids = vm.EmbeddingStore.load(p)
        assert ids[:3] == ["c0", "c1", "c2"] and len(st2) == n
        assert all(np.array_equal(a, b) for a, b in zip(ref, st2.topk(Q, 5)))
        st.close(); st2.close()


def test_async_device_conditional_exact_rescan(vm):
    """VM_FLAG_ASYNC never syncs: uncertified queries are re-done by the device-conditional exact scan."""
    import torch
    d, n, k = 64, 40000, 10
    rng = np.random.default_rng(11)
    X = rng.standard_normal((n, d)).astype(np.float32)
    X[500:900] = X[500]                                     # 400 exact ties: more than any candidate list holds
    Q = np.stack([X[500], rng.standard_normal(d).astype(np.float32), np.zeros(d, np.float32)] * 3)
    st = vm.EmbeddingStore(d, n, "f32")
    st.append(X)
    for flags in (vm.VM_FLAG_ASYNC, vm.VM_FLAG_ASYNC | vm.VM_FLAG_FORCE_TC, vm.VM_FLAG_ASYNC | vm.VM_FLAG_FORCE_SIMT):
        idx, score, count = st.topk_device(torch.from_numpy(Q).cuda(), k, sum_mode=vm.VM_SUM_NEUMAIER, flags=flags)
        torch.cuda.synchronize()
        _check(idx.cpu().numpy(), score.cpu().numpy(), count.cpu().numpy(), oracle.batch_similarities(Q, X, k), k)
        assert list(idx[0].cpu().numpy()) == list(range(500, 510))
    st.close()


def test_integration_md_ctypes_example(vm):
    """The stand-alone ctypes binding shown in INTEGRATION.md section 2, with library-owned store memory."""
    import ctypes as C
    lib = C.CDLL(vm._lib.LIB_PATH)
    lib.vm_last_error.restype = C.c_char_p
    n, d, k = 3000, 384, 10
    X = synth.synth_rows(41, 0, n, d).astype(np.float64)
    Qf = synth.synth_queries(42, 7, d, 41, n)
    store = C.c_void_p()
    assert lib.vm_store_create(C.byref(store), 0, d, 0, C.c_int64(n)) == 0, lib.vm_last_error()
    assert lib.vm_store_append(store, X.ctypes.data_as(C.c_void_p), 2, 0, C.c_int64(n), None, None) == 0, lib.vm_last_error()
    lib.vm_store_size.restype = C.c_int64
    assert lib.vm_store_size(store) == n
    q = Qf.astype(np.float64)
    idx = np.empty((len(q), k), np.int64); score = np.empty((len(q), k), np.float64); cnt = np.empty(len(q), np.int32)
    rc = lib.vm_topk(store, q.ctypes.data_as(C.c_void_p), 2, 0, len(q), k, C.c_double(-np.inf), 0, 1, 0,
                     idx.ctypes.data_as(C.c_void_p), score.ctypes.data_as(C.c_void_p), cnt.ctypes.data_as(C.c_void_p),
                     0, None, None)
    assert rc == 0, lib.vm_last_error()
    _check(idx, score, cnt, oracle.batch_similarities(Qf, X, k), k)
    # error convention: status code + message, nothing thrown across the ABI
    assert lib.vm_store_append(store, X.ctypes.data_as(C.c_void_p), 2, 0, C.c_int64(n), None, None) == vm._lib.VM_ERR_OVERFLOW
    assert b"capacity" in lib.vm_last_error()
    assert lib.vm_topk(store, q.ctypes.data_as(C.c_void_p), 2, 0, len(q), 1000, C.c_double(0), 0, 1, 0, None, None, None, 0, None, None) == vm._lib.VM_ERR_BADARG
    assert lib.vm_store_destroy(store) == 0


@pytest.mark.parametrize("scale", [1e-30, 1e-42, 1e25])
def test_extreme_magnitudes_stay_exact(vm, scale):
    """Rows far outside the fast scans' numeric range (fp32 denormals are flushed by the tensor cores,
    huge rows overflow fp32 accumulation) must not break exactness: the store flags them and the
    binary64 pass re-does the queries."""
    d, n, k = 128, 20000, 10
    rng = np.random.default_rng(7)
    X = rng.standard_normal((n, d)).astype(np.float32)
    X[100:140] = (X[100:140].astype(np.float64) * scale).astype(np.float32)   # 40 rows at an extreme scale
    Q = rng.standard_normal((6, d)).astype(np.float32)
    Q[0] = (X[120].astype(np.float64) / scale).astype(np.float32)             # points at one of them
    st = vm.EmbeddingStore(d, n, "f32")
    st.append(X)
    ref = oracle.batch_similarities(Q, X, k)
    for flags in (0, vm.VM_FLAG_FORCE_TC, vm.VM_FLAG_FORCE_SIMT):
        idx, score, count = st.topk(Q, k, sum_mode=vm.VM_SUM_NEUMAIER, flags=flags)
        _check(idx, score, count, ref, k)
    assert st.last_stats.uncertified == len(Q)          # every query went through the exact pass
    assert ref[0][0][0] == 120
    st.close()


def test_non_finite_rows_are_skipped(vm):
    d, n = 32, 500
    X = synth.synth_rows(5, 0, n, d)
    Xbad = X.copy(); Xbad[10, 3] = np.nan; Xbad[20, 0] = np.inf
    st = vm.EmbeddingStore(d, n, "f32")
    st.append(Xbad)
    ok = np.ones(n, np.uint8); ok[[10, 20]] = 0
    Q = X[[10, 20, 30]]
    _check(*st.topk(Q, 5, sum_mode=vm.VM_SUM_NEUMAIER), oracle.batch_similarities(Q, X, 5, row_ok=ok), 5)
    st.close()


def test_randomised_differential(vm):
    """Fuzz: random shapes / dtypes / k / duplicate structure, every scan path against the oracle."""
    rng = np.random.default_rng(20261018)
    for case in range(36):
        n = int(rng.choice([1, 2, 31, 128, 129, 700, 3001, 9000]))
        d = int(rng.choice([8, 24, 100, 384, 520]))
        nq = int(rng.choice([1, 3, 8, 9, 40, 64, 70]))
        k = int(rng.choice([1, 3, 10, 16, 24, 33]))
        dtype = "bf16" if case % 3 == 0 else "f32"
        vals = rng.integers(-2, 3, size=(n, d)).astype(np.float32) if case % 4 == 0 else rng.standard_normal((n, d)).astype(np.float32)
        X = _quantise(vals, dtype)
        if n > 40:
            X[rng.integers(0, n, 6)] = X[7]                # duplicates -> exact ties
            X[rng.integers(0, n, 2)] = 0.0
        Q = rng.standard_normal((nq, d)).astype(np.float32)
        if n > 8:
            Q[0] = X[7]
        st = vm.EmbeddingStore(d, n, dtype)
        st.append(X)
        ref = oracle.batch_similarities(Q, X, k)
        for flags in (0, vm.VM_FLAG_FORCE_TC, vm.VM_FLAG_FORCE_SIMT, vm.VM_FLAG_FORCE_EXACT):
            idx, score, count = st.topk(Q, k, sum_mode=vm.VM_SUM_NEUMAIER, flags=flags)
            try:
                _check(idx, score, count, ref, k)
            except AssertionError as e:
                raise AssertionError(f"case {case}: n={n} d={d} nq={nq} k={k} {dtype} flags={flags}: {e}")
        st.close()


def test_boundaries_empty_store_large_k_no_queries(vm):
    d = 32
    st = vm.EmbeddingStore(d, 100, "f32")
    Q = synth.synth_queries(1, 3, d, 2, 100)
    idx, score, count = st.topk(Q, 5)                                # empty store -> nothing, no error
    assert (count == 0).all() and (idx == -1).all()
    X = synth.synth_rows(2, 0, 70, d)
    st.append(X)
    idx, score, count = st.topk(Q[:0], 5)                            # no queries
    assert idx.shape == (0, 5)
    for k in (49, 64):                                               # beyond the fast scans' candidate lists: exact scan
        idx, score, count = st.topk(Q, k, sum_mode=vm.VM_SUM_NEUMAIER)
        _check(idx, score, count, oracle.batch_similarities(Q, X, k), k)
        assert st.last_stats.scan_kernel == 0
    with pytest.raises(vm.VidmemError):
        st.topk(Q, 65)
    with pytest.raises(ValueError):
        st.topk(np.zeros((2, d + 1), np.float32), 3)
    st.close()


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_near_duplicate_cluster_is_settled_from_the_band(vm, dtype):
    """A cluster of near-identical rows (video chunks of a static scene) puts more rows inside the scan's
    error band than the rescoring kernel takes at first: the scan has kept the COMPLETE band in the query's union
    buffer, so the same kernel rescores all of it -- no second scan, no binary64 scan of every row."""
    import torch
    d, n, k = 384, 120000, 10
    rng = np.random.default_rng(23)
    X = rng.standard_normal((n, d)).astype(np.float32)
    base = rng.standard_normal(d).astype(np.float32)
    X[5000:5300] = base + 1e-3 * rng.standard_normal((300, d)).astype(np.float32)   # cosines within ~1e-6 of each other
    X = _quantise(X, dtype)
    Q = np.stack([base, rng.standard_normal(d).astype(np.float32), X[77]])
    st = vm.EmbeddingStore(d, n, dtype)
    st.append(X)
    ref = oracle.topk_blocked(Q, X, k, slack=400)
    with torch.cuda.stream(torch.cuda.Stream()):                      # non-default stream: plain (non-graph) sync path, stats are read back
        idx, score, count = st.topk(Q, k, sum_mode=vm.VM_SUM_NEUMAIER)
    assert np.array_equal(idx, ref[0]) and np.array_equal(score, ref[1])
    c = st.counters(reset=True)
    assert st.last_stats.scan_kernel == 2 and st.last_stats.uncertified == 0      # nothing was left for a fallback pass
    assert c["band_settled"] >= 1 and c["collect_settled"] == 0 and c["full_rescans"] == 0 and c["bound_violations"] == 0
    for _ in range(20):                                               # default stream: CUDA-graph replay after 16 calls, same answer
        idx2, score2, _ = st.topk(Q, k, sum_mode=vm.VM_SUM_NEUMAIER)
    assert np.array_equal(idx2, idx) and np.array_equal(score2, score)
    qd = torch.from_numpy(Q).cuda()                                   # device path, no host round trip at all
    i4, s4, _ = st.topk_device(qd, k, sum_mode=vm.VM_SUM_NEUMAIER, flags=vm.VM_FLAG_ASYNC)
    torch.cuda.synchronize()
    assert np.array_equal(i4.cpu().numpy(), idx) and np.array_equal(s4.cpu().numpy(), score)
    # a band of 1500 rows is still settled by the rescoring kernel itself (it takes up to 2048)
    X0 = X.copy(); X0[30000:31200] = X0[5000:5001] + 0.0
    st0 = vm.EmbeddingStore(d, n, dtype); st0.append(X0)
    ref0 = oracle.topk_blocked(Q[:1], X0, k, slack=2000)
    with torch.cuda.stream(torch.cuda.Stream()):
        i0, s0, _ = st0.topk(Q[:1], k, sum_mode=vm.VM_SUM_NEUMAIER)
    c0 = st0.counters()
    assert np.array_equal(i0, ref0[0]) and np.array_equal(s0, ref0[1])
    assert c0["band_settled"] == 1 and c0["collect_settled"] == 0 and c0["full_rescans"] == 0
    st0.close()
    # a band of 2800 rows: beyond what the rescoring kernel takes itself (2048), within the collect pass (4096)
    X1 = X.copy(); X1[30000:32500] = X1[5000:5001] + 0.0
    st1 = vm.EmbeddingStore(d, n, dtype); st1.append(X1)
    ref1 = oracle.topk_blocked(Q[:1], X1, k, slack=3400)
    with torch.cuda.stream(torch.cuda.Stream()):
        i5, s5, _ = st1.topk(Q[:1], k, sum_mode=vm.VM_SUM_NEUMAIER)
    c1 = st1.counters()
    assert np.array_equal(i5, ref1[0]) and np.array_equal(s5, ref1[1])
    assert c1["collect_settled"] == 1 and c1["full_rescans"] == 0 and st1.last_stats.uncertified == 1
    i6, s6, _ = st1.topk_device(qd[:1].contiguous(), k, sum_mode=vm.VM_SUM_NEUMAIER, flags=vm.VM_FLAG_ASYNC)   # async: straight to the exact scan
    torch.cuda.synchronize()
    assert np.array_equal(i6.cpu().numpy(), ref1[0]) and np.array_equal(s6.cpu().numpy(), ref1[1])
    assert st1.counters()["full_rescans"] == 1
    # more near-ties than the union and collect buffers hold -> the binary64 scan takes over, still exact
    X2 = X.copy(); X2[20000:25000] = X2[5000]
    st2 = vm.EmbeddingStore(d, n, dtype); st2.append(X2)
    ref2 = oracle.topk_blocked(Q[:1], X2, k, slack=6000)
    with torch.cuda.stream(torch.cuda.Stream()):
        i3, s3, _ = st2.topk(Q[:1], k, sum_mode=vm.VM_SUM_NEUMAIER)
    assert np.array_equal(i3, ref2[0]) and np.array_equal(s3, ref2[1]) and st2.last_stats.full_rescans == 1
    st.close(); st1.close(); st2.close()


@pytest.mark.parametrize("dtype,flags", [("f32", 0), ("bf16", 0), ("bf16", 32), ("bf16", 64)])
def test_clustered_store_single_pass(vm, dtype, flags):
    """Unit-norm Gaussian-mixture store (64 near-duplicates per centre, stored consecutively like the chunks of one
    scene; values not representable in bf16 / tf32) and queries that are perturbed members: every top-10 sits inside a
    near-duplicate cluster.  Exact against the oracle, settled in ONE scan, and the scan's error bound holds on every
    rescored candidate.  flags=32 (VM_FLAG_NO_SPLIT): the single-term bf16 query, whose band is 60x wider than that of
    flags=64 (VM_FLAG_SPLIT: hi + lo query terms; the default up to 48 queries per batch)."""
    import torch
    d, n, nq, k, per, rho = 384, 131072 + 777, 64, 10, 64, 0.2
    rng = np.random.default_rng(99)
    cent = rng.standard_normal((n // per + 1, d)).astype(np.float32)
    cent /= np.linalg.norm(cent, axis=1, keepdims=True)
    X = cent[np.arange(n) // per] + (rho / np.sqrt(d)) * rng.standard_normal((n, d)).astype(np.float32)
    X /= np.linalg.norm(X, axis=1, keepdims=True)
    X = _quantise(X.astype(np.float32), dtype)
    pick = rng.integers(0, n, nq)
    Q = (X[pick] + (rho / np.sqrt(d)) * rng.standard_normal((nq, d))).astype(np.float32)
    st = vm.EmbeddingStore(d, n, dtype)
    st.append(X)
    ref = oracle.topk_blocked(Q, X, k, slack=120)
    idx, score, count = st.topk_device(torch.from_numpy(Q).cuda(), k, sum_mode=vm.VM_SUM_NEUMAIER, flags=vm.VM_FLAG_ASYNC | flags)
    torch.cuda.synchronize()
    assert np.array_equal(idx.cpu().numpy(), ref[0]) and np.array_equal(score.cpu().numpy(), ref[1])
    c = st.counters()
    assert c["queries"] == nq and c["bound_violations"] == 0 and c["full_rescans"] == 0 and c["collect_settled"] == 0
    assert c["uncertified"] == c["band_settled"]
    if flags == 64:
        assert c["uncertified"] <= 2                                  # split query: the band is ~1e-4 wide
    if dtype == "bf16":                                               # batches of <= 48 queries take the split path by default
        i2, s2, _ = st.topk_device(torch.from_numpy(Q[:40]).cuda(), k, sum_mode=vm.VM_SUM_NEUMAIER, flags=vm.VM_FLAG_ASYNC)
        torch.cuda.synchronize()
        assert np.array_equal(i2.cpu().numpy(), ref[0][:40]) and np.array_equal(s2.cpu().numpy(), ref[1][:40])
    st.close()


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
@pytest.mark.parametrize("n", [127, 128, 129, 4095, 4097, 9472, 9473])
def test_small_store_dump_mode_boundaries(vm, dtype, n):
    """Stores of <= 9472 rows take the tcgen05 scan's dump mode (every row's key is ranked, no per-CTA lists);
    9473 rows is the first size on the list path.  Skipped, zero and duplicate rows, partial last tile."""
    d, nq, k = 96, 19, 10
    rng = np.random.default_rng(n)
    X = _quantise(rng.standard_normal((n, d)).astype(np.float32), dtype)
    X[n - 1] = X[1]; X[50] = X[1]; X[7] = 0.0
    Q = rng.standard_normal((nq, d)).astype(np.float32)
    Q[0] = X[1]; Q[1] = 0.0
    st = vm.EmbeddingStore(d, n, dtype)
    st.append(X)
    st.invalidate([2, n - 2])
    ok = np.ones(n, np.uint8); ok[[2, n - 2]] = 0
    ref = oracle.batch_similarities(Q, X, k, row_ok=ok)
    idx, score, count = st.topk(Q, k, sum_mode=vm.VM_SUM_NEUMAIER)
    assert st.last_stats.scan_kernel == 2 and (st.last_stats.scan_variant == 1) == (n <= 9472)
    _check(idx, score, count, ref, k)
    assert list(idx[0, :3]) == [1, 50, n - 1]                       # three-way tie -> lowest rows first
    st.close()


def test_small_store_near_ties_band_after_dump(vm):
    """More near-identical rows than the candidate list holds, in a store small enough for dump mode: every row's
    key is in the dump, so the rescoring kernel settles the query from the band itself."""
    import torch
    d, n, k = 384, 6000, 10
    rng = np.random.default_rng(5)
    X = rng.standard_normal((n, d)).astype(np.float32)
    base = rng.standard_normal(d).astype(np.float32)
    X[1000:1200] = base + 1e-3 * rng.standard_normal((200, d)).astype(np.float32)
    Q = np.stack([base] + [rng.standard_normal(d).astype(np.float32) for _ in range(11)])
    st = vm.EmbeddingStore(d, n, "f32")
    st.append(X)
    ref = oracle.batch_similarities(Q, X, k)
    with torch.cuda.stream(torch.cuda.Stream()):
        idx, score, count = st.topk(Q, k, sum_mode=vm.VM_SUM_NEUMAIER)
    _check(idx, score, count, ref, k)
    c = st.counters()
    assert st.last_stats.scan_kernel == 2 and st.last_stats.scan_variant == 1 and st.last_stats.uncertified == 0
    assert c["band_settled"] >= 1 and c["collect_settled"] == 0 and c["full_rescans"] == 0
    st.close()


def test_avg_scan_ms_covers_every_timed_call(vm):
    import torch
    n, d = 20000, 64
    X = np.random.default_rng(1).standard_normal((n, d)).astype(np.float32)
    st = vm.EmbeddingStore(d, n, "f32"); st.append(X)
    q = torch.from_numpy(X[:16].copy()).cuda()
    with pytest.raises(vm.VidmemError):
        st.avg_scan_ms()                                             # nothing timed yet
    for _ in range(5):
        st.topk_device(q, 10, flags=vm.VM_FLAG_ASYNC | vm.VM_FLAG_TIMING)
    ms, calls = st.avg_scan_ms()
    assert calls == 5 and 0.0 < ms < 50.0
    for _ in range(70):                                              # more than the ring holds: the last 64 are averaged
        st.topk_device(q, 10, flags=vm.VM_FLAG_ASYNC | vm.VM_FLAG_TIMING)
    ms2, calls2 = st.avg_scan_ms()
    assert calls2 == 64 and 0.0 < ms2 < 50.0 and abs(st.last_scan_ms() - ms2) < 10 * ms2
    with pytest.raises(vm.VidmemError):
        st.avg_scan_ms()                                             # already consumed
    st.close()


def test_threshold_warp_scan_vs_blocked_oracle(vm):
    """bf16, 1.3 M rows, 40 queries: long enough for the threshold-warp variant of the tcgen05 scan (shared bound
    from every CTA's running maxima).  All queries checked against the blocked oracle, incl. planted duplicates."""
    import torch
    n, d, nq, k = 1_300_000, 384, 40, 10
    st = vm.EmbeddingStore(d, n, "bf16")
    st.synth_fill(11, n, dup_period=50_000)                          # planted exact duplicates -> cross-CTA ties
    st.set_size(n)
    X = st.rows[:n, :d].float().cpu().numpy()
    Q = synth.synth_queries(12, nq, d, 11, n)
    Q[0] = X[123_456]; Q[1] = X[n - 1]
    with torch.cuda.stream(torch.cuda.Stream()):
        idx, score, count = st.topk(Q, k, sum_mode=vm.VM_SUM_NEUMAIER)
    assert st.last_stats.scan_kernel == 2 and st.last_stats.scan_variant == 2
    bi, bs, bc = oracle.topk_blocked(Q, X, k, slack=64)
    assert np.array_equal(idx, bi) and np.array_equal(score, bs) and (count == k).all()
    st.close()
