import sys, time, torch, numpy as np
sys.path.insert(0, ".")
import vidmem_b200 as vm
from vidmem_b200 import dedup
from vidmem_b200.store import EmbeddingStore
for n, d in ((131072, 768), (524288, 768)):
    st = EmbeddingStore(d, n, "bf16")
    st.synth_fill(4, n, dup_period=100)
    torch.cuda.synchronize()
    x = st.rows[:n]
    for it in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        i, j, s = dedup.pairs_above(x, 0.9, cap=1 << 22)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    pairs = n * (n - 1) / 2
    print(f"n={n} d={d}: {ms:.2f} ms  hits={len(i)}  {pairs/ms*1e3:.3e} pairs/s  {pairs*2*d/ms*1e3/1e12:.1f} TFLOP/s (triangle)")
    st.close()
