"""Config C1 (the reference's own size: ~5K x 384 store, 30 queries, top-10): per-call latency through the
public host-buffer API, and the CPU oracle port beside it."""
import sys, time, numpy as np, torch
sys.path.insert(0, ".")
import vidmem_b200 as vm
from oracle import oracle, synth
n, d, q, k = 5000, 384, 30, 10
X = synth.synth_rows(1, 0, n, d); Q = synth.synth_queries(1001, q, d, 1, n)
st = vm.EmbeddingStore(d, n, "f32"); st.append(X)
for flags, name in ((0, "default"), (vm.VM_FLAG_FORCE_SIMT, "simt"), (vm.VM_FLAG_FORCE_TC, "tcgen05")):
    for _ in range(20): st.topk(Q, k, flags=flags)
    t0 = time.perf_counter()
    for _ in range(200): st.topk(Q, k, flags=flags)
    dt = (time.perf_counter() - t0) / 200
    print(f"C1 {name}: {dt*1e6:.1f} us per 30-query call ({q/dt:.0f} q/s), scan kernel {st.last_stats.scan_kernel}, launches {st.last_stats.scan_launches}")
t0 = time.perf_counter(); oracle.batch_similarities(Q, X, k); print(f"oracle port (all cores): {(time.perf_counter()-t0)*1e3:.1f} ms")
