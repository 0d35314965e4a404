"""GPU parity of the all-pairs threshold scorer against the sklearn-semantics oracle and the golden
pair set produced by sklearn itself (tests/golden/prune.npz)."""
import os

import numpy as np
import pytest

from oracle import oracle, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dd():
    import torch
    assert torch.cuda.is_available()
    from vidmem_b200 import dedup
    return dedup


def _check(dd, E, thr, dtype):
    import torch
    x = torch.from_numpy(E).cuda().to(dtype)
    Eq = x.float().cpu().numpy()                       # the stored (rounded) values are what both sides score
    i, j, s = dd.pairs_above(x, thr)
    oi, oj, os_ = oracle.pairs_above(Eq, thr)
    assert list(zip(i.tolist(), j.tolist())) == list(zip(oi.tolist(), oj.tolist()))
    np.testing.assert_allclose(s, os_, rtol=1e-3 if dtype == torch.float32 else 2e-4, atol=1e-6)
    return len(oi)


def test_golden_sklearn_pair_set(dd, golden_dir):
    import torch
    g = np.load(os.path.join(golden_dir, "prune.npz"))
    E = synth.synth_rows(int(g["pairs_seed"]), 0, int(g["pairs_n"]), int(g["pairs_d"]), int(g["pairs_dup"]))
    E[50] = 0.0
    for thr in (0.8, 0.9):
        t = int(thr * 10)
        for dtype in (torch.bfloat16, torch.float32):     # synthetic values are exact in bf16 and tf32
            i, j, s = dd.pairs_above(torch.from_numpy(E).cuda().to(dtype), thr)
            assert np.array_equal(i, g[f"pairs_i_{t}"]) and np.array_equal(j, g[f"pairs_j_{t}"])
            np.testing.assert_allclose(s, g[f"pairs_s_{t}"], rtol=1e-5)


@pytest.mark.parametrize("n,d,dup", [(2, 8, 1), (100, 64, 3), (129, 768, 4), (257, 384, 5), (1000, 768, 7),
                                     (2500, 100, 9), (4500, 768, 50)])
def test_vs_oracle_bf16(dd, n, d, dup):
    import torch
    E = synth.synth_rows(1000 + n, 0, n, d, dup_period=dup)
    if n > 10:
        E[7] = 0.0
        E[n - 1] = E[3]
    found = _check(dd, E, 0.9, torch.bfloat16)
    assert found > 0 or n < 10


def test_vs_oracle_fp32_and_general_values(dd):
    import torch
    rng = np.random.default_rng(3)
    base = rng.standard_normal((40, 256)).astype(np.float32)
    E = np.concatenate([base + 0.15 * rng.standard_normal((40, 256)).astype(np.float32) for _ in range(20)])
    _check(dd, E, 0.9, torch.float32)
    _check(dd, E, 0.97, torch.bfloat16)


def test_edge_semantics(dd):
    import torch
    x = torch.from_numpy(synth.synth_rows(5, 0, 1, 64)).cuda().to(torch.bfloat16)
    i, j, s = dd.pairs_above(x, 0.5)
    assert len(i) == 0                                   # n <= 1 -> no pairs (prune.py:73-74)
    E = synth.synth_rows(6, 0, 300, 64, dup_period=2)
    with pytest.raises(Exception):
        dd.pairs_above(torch.from_numpy(E).cuda().to(torch.bfloat16), 0.5, cap=4)   # overflow is loud
    # splitting the tile grid across "ranks" partitions the pair set
    x = torch.from_numpy(synth.synth_rows(7, 0, 3000, 128, dup_period=6)).cuda().to(torch.bfloat16)
    whole = dd.pairs_above(x, 0.9)
    parts = [dd.pairs_above(x, 0.9, part=p, nparts=3) for p in range(3)]
    got = sorted(sum([list(zip(a.tolist(), b.tolist())) for a, b, _ in parts], []))
    assert got == list(zip(whole[0].tolist(), whole[1].tolist())) and len(got) > 0


def test_two_streams_do_not_share_scratch(dd):
    """Two calls enqueued back to back on different streams (different inputs, nothing synchronised in between): each
    gets its own pair set -- the library keeps its scratch per (device, stream)."""
    import ctypes as C
    import torch
    import vidmem_b200 as vm
    lib = vm._lib.load()
    dev = torch.device("cuda", 0)
    cap = 1 << 16
    jobs = []
    for seed, n in ((11, 6000), (12, 9000)):
        E = synth.synth_rows(seed, 0, n, 256, dup_period=7)
        x = torch.from_numpy(E).to(dev).to(torch.bfloat16)
        out = (torch.empty(cap, dtype=torch.int64, device=dev), torch.empty(cap, dtype=torch.int64, device=dev),
               torch.empty(cap, dtype=torch.float32, device=dev), torch.zeros(1, dtype=torch.int64, device=dev))
        jobs.append((E, x, out, torch.cuda.Stream(dev)))
    torch.cuda.synchronize()
    for _ in range(3):                                    # interleaved, asynchronous
        for E, x, out, st in jobs:
            vm._lib.check(lib.vm_pairs_above(0, x.data_ptr(), vm.VM_BF16, x.shape[0], x.shape[1], C.c_float(0.9), cap, out[0].data_ptr(),
                                             out[1].data_ptr(), out[2].data_ptr(), out[3].data_ptr(), 0, 1, vm.VM_FLAG_ASYNC, st.cuda_stream))
    torch.cuda.synchronize()
    for E, x, out, st in jobs:
        m = int(out[3].item())
        got = sorted(zip(out[0][:m].tolist(), out[1][:m].tolist()))
        oi, oj, _ = oracle.pairs_above(E, 0.9)
        assert m > 50 and got == list(zip(oi.tolist(), oj.tolist()))


@pytest.mark.parametrize("case", range(12))
def test_random_pairs_near_the_threshold(dd, case):
    """Seeded random stress: clusters whose internal cosine is drawn AROUND the threshold (the hard regime for the
    epilogue's band and the binary64 re-decision), random sizes / dimensions / dtypes; the pair set must equal the
    oracle's on the stored values."""
    import torch
    rng = np.random.default_rng(500 + case)
    n = int(rng.integers(300, 14_000))
    d = int(rng.choice([64, 96, 256, 384, 768]))
    thr = float(rng.choice([0.8, 0.9, 0.95]))
    per = int(rng.choice([2, 3, 8, 40]))
    # cosine between cluster members ~ 1 / (1 + s^2): s chosen so that it straddles thr
    s_mid = (1.0 / thr - 1.0) ** 0.5
    s = s_mid * (1.0 + rng.uniform(-0.3, 0.3, size=(n // per + 1, 1)))
    centres = rng.standard_normal((n // per + 1, d)) / d ** 0.5
    E = (centres[np.arange(n) // per] + s[np.arange(n) // per] * rng.standard_normal((n, d)) / d ** 0.5 / 2 ** 0.5).astype(np.float32)
    E[5] = 0.0
    E[n - 1] = E[0]
    dtype = torch.bfloat16 if case % 2 == 0 else torch.float32
    found = _check(dd, E, thr, dtype)
    assert found > 0

