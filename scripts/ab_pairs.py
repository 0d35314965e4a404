"""Same-box A/B of all-pairs kernel builds: python scripts/ab_pairs.py ROWS ITERS   (library chosen with VIDMEM_LIB).
One line: library, ms per launch (CUDA events, each launch), hits, pair checksum, SM clock / power sampled during the last launch."""
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C

import torch

import vidmem_b200 as vm
from vidmem_b200.store import EmbeddingStore

n, iters, d, thr = int(sys.argv[1]), int(sys.argv[2]), 768, 0.9
st = EmbeddingStore(d, n, "bf16")
st.synth_fill(4, n, dup_period=100)
x = st.rows[:n]
cap = 1 << 22
dev = x.device
oi = torch.empty((cap,), dtype=torch.int64, device=dev); oj = torch.empty_like(oi)
os_ = torch.empty((cap,), dtype=torch.float32, device=dev); cnt = torch.zeros((1,), dtype=torch.int64, device=dev)
lib = vm._lib.load()
samples, stop = [], False


def poll():
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    while not stop:
        samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0))
        time.sleep(0.02)


torch.cuda.synchronize()
ms = []
for it in range(iters):
    if it == iters - 1:
        th = threading.Thread(target=poll); th.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    vm._lib.check(lib.vm_pairs_above(0, x.data_ptr(), vm.VM_BF16, n, d, C.c_float(thr), cap, oi.data_ptr(), oj.data_ptr(), os_.data_ptr(),
                                     cnt.data_ptr(), 0, 1, 0, torch.cuda.current_stream(dev).cuda_stream))
    e1.record(); torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1))
stop = True; th.join()
m = int(cnt.item())
h = int(((oi[:m] * 1_000_003 + oj[:m]) % 2_147_483_629).sum().item())
clk = sorted(s[0] for s in samples)[len(samples) // 2] if samples else None
pw = sorted(s[1] for s in samples)[len(samples) // 2] if samples else None
pairs = n * (n - 1) / 2
print(f"{os.path.basename(vm._lib.LIB_PATH)} n={n} ms={[round(v, 1) for v in ms]} hits={m} checksum={h} "
      f"TFLOPs(last)={pairs * 2 * d / ms[-1] * 1e3 / 1e12:.0f} sm_mhz={clk} power_w={pw}")
st.close()
