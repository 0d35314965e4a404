"""Where the host-buffer call's time goes (C1 and C2 shapes): the Python wrapper, the bare C call, and the
device work inside it (graph replay) -- to decide whether the wrapper is worth trimming."""
import ctypes as C, sys, time, numpy as np, torch
sys.path.insert(0, ".")
import vidmem_b200 as vm
from vidmem_b200 import _lib as L
from oracle import synth

def bench(fn, n=300, warm=40):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    return (time.perf_counter() - t0) / n * 1e6

for name, n, nq in (("C1", 5000, 30), ("C2", 1_000_000, 64)):
    d, k = 384, 10
    st = vm.EmbeddingStore(d, n, "f32"); st.synth_fill(2, n); st.set_size(n)
    Q = synth.synth_queries(1001, nq, d, 2, n).astype(np.float32)
    lib = st.lib
    idx = np.full((nq, k), -1, np.int64); score = np.zeros((nq, k)); count = np.zeros(nq, np.int32)
    args = (st._h, Q.ctypes.data, L.VM_F32, L.VM_MEM_HOST, nq, k, float("-inf"), 0, L.DEFAULT_SUM_MODE, 0,
            idx.ctypes.data, score.ctypes.data, count.ctypes.data, L.VM_MEM_HOST, None, None)
    t_py = bench(lambda: st.topk(Q, k))
    t_c = bench(lambda: lib.vm_topk(*args))
    qd = torch.from_numpy(Q).cuda()
    out = (torch.empty((nq, k), dtype=torch.int64, device="cuda"), torch.empty((nq, k), dtype=torch.float64, device="cuda"),
           torch.empty((nq,), dtype=torch.int32, device="cuda"))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(20): st.topk_device(qd, k, out=out, flags=vm.VM_FLAG_ASYNC)
    torch.cuda.synchronize(); e0.record()
    for _ in range(200): st.topk_device(qd, k, out=out, flags=vm.VM_FLAG_ASYNC)
    e1.record(); torch.cuda.synchronize()
    t_dev = e0.elapsed_time(e1) / 200 * 1e3
    print(f"{name}: python wrapper {t_py:.1f} us, bare C call {t_c:.1f} us, device-resident step {t_dev:.1f} us "
          f"-> wrapper adds {t_py - t_c:.1f} us, host path adds {t_c - t_dev:.1f} us over the device work")
    st.close()
