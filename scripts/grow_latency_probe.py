"""How long one growth step of a growable store takes (vm_store_reserve: cuMemCreate + cuMemMap + cuMemSetAccess of the new
chunks), and how long the Python side needs to rebuild its tensor views -- on a fresh device and after another large
allocation has been freed (the state a long-running process is in)."""
import ctypes as C, sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import vidmem_b200 as vm
from vidmem_b200 import _lib as L

def probe(tag, dt, rows0, step_rows, n):
    st = vm.EmbeddingStore(384, rows0, dt, max_capacity=rows0 + (n + 2) * step_rows)
    st.synth_fill(5, rows0); st.set_size(rows0); torch.cuda.synchronize()
    lib = st.lib
    t_c, t_py = [], []
    for i in range(n):
        cap = st.capacity + step_rows
        torch.cuda.synchronize()
        t0 = time.perf_counter(); L.check(lib.vm_store_reserve(st._h, cap)); t1 = time.perf_counter()
        st._refresh_views(); t2 = time.perf_counter()
        t_c.append((t1 - t0) * 1e3); t_py.append((t2 - t1) * 1e3)
    t_c, t_py = np.array(t_c), np.array(t_py)
    print(f"{tag}: reserve (+{step_rows} rows) ms p50 {np.median(t_c):.3f} max {t_c.max():.3f} first5 {np.round(t_c[:5], 3)} | views ms p50 {np.median(t_py):.3f} max {t_py.max():.3f}")
    st.close()

probe("fresh f32", "f32", 2_000_000, 170_000, 30)
x = torch.empty(60 * (1 << 30), dtype=torch.uint8, device="cuda"); x.zero_(); torch.cuda.synchronize(); del x; torch.cuda.empty_cache()
probe("after a freed 60 GB torch block, f32", "f32", 2_000_000, 170_000, 30)
big = vm.EmbeddingStore(384, 10_000_000, "f32", max_capacity=12_000_000); big.synth_fill(5, 10_000_000); big.set_size(10_000_000); torch.cuda.synchronize(); big.close()
probe("after a closed 15 GB growable store, bf16", "bf16", 10_000_000, 340_000, 30)
