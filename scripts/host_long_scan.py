"""Long scans through the host-buffer call: is the extra time (vs the device-resident loop) on the GPU (scan kernel
slower after an idle gap) or on the host (wait / wake-up / copies)?"""
import sys, time, numpy as np, torch
sys.path.insert(0, ".")
import vidmem_b200 as vm
from oracle import synth
n, d, nq, k = 12_500_000, 384, 64, 10
st = vm.EmbeddingStore(d, n, "bf16"); st.synth_fill(3, n); st.set_size(n)
Q = synth.synth_queries(3003, nq, d, 3, n).astype(np.float32)
qd = torch.from_numpy(Q).cuda()
out = (torch.empty((nq, k), dtype=torch.int64, device="cuda"), torch.empty((nq, k), dtype=torch.float64, device="cuda"),
       torch.empty((nq,), dtype=torch.int32, device="cuda"))
for _ in range(10): st.topk_device(qd, k, out=out, flags=vm.VM_FLAG_ASYNC | vm.VM_FLAG_TIMING)
torch.cuda.synchronize(); st.avg_scan_ms()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(40): st.topk_device(qd, k, out=out, flags=vm.VM_FLAG_ASYNC | vm.VM_FLAG_TIMING)
e1.record(); torch.cuda.synchronize()
print(f"device loop: {e0.elapsed_time(e1)/40:.4f} ms/step, scan {st.avg_scan_ms()[0]:.4f} ms")
for name, flags in (("host plain+timing", vm.VM_FLAG_TIMING), ("host default (graph after 16 calls)", 0)):
    for _ in range(20): st.topk(Q, k, flags=flags)
    if flags: st.avg_scan_ms()
    t0 = time.perf_counter()
    for _ in range(40): st.topk(Q, k, flags=flags)
    dt = (time.perf_counter() - t0) / 40 * 1e3
    print(f"{name}: {dt:.4f} ms/call" + (f", scan {st.avg_scan_ms()[0]:.4f} ms" if flags else ""))
# same host calls with a busy GPU in between is not possible (synchronous API); instead: host call right after a dummy kernel
x = torch.empty(64 << 20, device="cuda")
t0 = time.perf_counter()
for _ in range(40):
    x.zero_()                      # ~0.04 ms of GPU work enqueued just before the call: no idle gap before the scan
    st.topk(Q, k, flags=vm.VM_FLAG_TIMING)
dt = (time.perf_counter() - t0) / 40 * 1e3
print(f"host plain+timing with a kernel enqueued just before: {dt:.4f} ms/call, scan {st.avg_scan_ms()[0]:.4f} ms")
