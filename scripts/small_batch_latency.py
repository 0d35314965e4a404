"""Few queries over small/medium stores: CUDA-core scan vs the tcgen05 scan (dump mode / seeded lists).
Decides the kernel-selection rule for nq <= 8."""
import sys, time, numpy as np, torch
sys.path.insert(0, ".")
import vidmem_b200 as vm
from oracle import synth
d, k = 384, 10
for n in (5000, 9000, 20000, 60000):
    st = vm.EmbeddingStore(d, n, "f32"); st.synth_fill(1, n); st.set_size(n)
    for nq in (1, 4, 8):
        Q = synth.synth_queries(1001, nq, d, 1, n).astype(np.float32)
        qd = torch.from_numpy(Q).cuda()
        out = (torch.empty((nq, k), dtype=torch.int64, device="cuda"), torch.empty((nq, k), dtype=torch.float64, device="cuda"),
               torch.empty((nq,), dtype=torch.int32, device="cuda"))
        res = []
        for flags, name in ((vm.VM_FLAG_FORCE_SIMT, "simt"), (vm.VM_FLAG_FORCE_TC, "tc")):
            f = flags | vm.VM_FLAG_ASYNC
            for _ in range(30): st.topk_device(qd, k, out=out, flags=f)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            for _ in range(300): st.topk_device(qd, k, out=out, flags=f)
            e1.record(); torch.cuda.synchronize()
            res.append(f"{name} {e0.elapsed_time(e1) / 300 * 1e3:.1f} us (variant {st.last_stats.scan_variant})")
        print(f"n={n} nq={nq}: " + ", ".join(res))
    st.close()
