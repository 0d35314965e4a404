"""How much GPU activity just before a long scan removes the idle-exit penalty of the scan kernel?"""
import sys, time, numpy as np, torch
sys.path.insert(0, ".")
import vidmem_b200 as vm
from oracle import synth
n, d, nq, k = 12_500_000, 384, 64, 10
st = vm.EmbeddingStore(d, n, "bf16"); st.synth_fill(3, n); st.set_size(n)
Q = synth.synth_queries(3003, nq, d, 3, n).astype(np.float32)
x = torch.empty(64 << 20, device="cuda")          # 256 MB of fp32
rows = st.rows
for name, pre in (("nothing", None), ("zero 4 MB", lambda: x[:1 << 20].zero_()), ("zero 16 MB", lambda: x[:4 << 20].zero_()),
                  ("zero 64 MB", lambda: x[:16 << 20].zero_()), ("read-sum 32 MB of the store", lambda: rows[:43690].sum()),
                  ("zero 256 MB", lambda: x.zero_()), ("nothing", None)):
    for _ in range(10): st.topk(Q, k, flags=vm.VM_FLAG_TIMING)
    st.avg_scan_ms()
    t0 = time.perf_counter()
    for _ in range(30):
        if pre: pre()
        st.topk(Q, k, flags=vm.VM_FLAG_TIMING)
    dt = (time.perf_counter() - t0) / 30 * 1e3
    print(f"{name}: {dt:.4f} ms/call, scan {st.avg_scan_ms()[0]:.4f} ms")
