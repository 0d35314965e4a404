"""Same-box A/B of scan kernels across library builds (only the C-ABI symbols every build has).
Usage: python scripts/ab_scan.py LIB ROWS DTYPE NQ [FLAGS] [ITERS]   -> one line: lib, scan ms (events around the scan), step ms"""
import ctypes as C
import sys

import torch

lib_path, rows, dt, nq = sys.argv[1], int(sys.argv[2]), sys.argv[3], int(sys.argv[4])
flags = int(sys.argv[5]) if len(sys.argv) > 5 else 0
iters = int(sys.argv[6]) if len(sys.argv) > 6 else 20
dim, k = 384, 10
L = C.CDLL(lib_path)
L.vm_last_error.restype = C.c_char_p
code = 0 if dt == "f32" else 1
tdt = torch.float32 if dt == "f32" else torch.bfloat16
X = torch.empty((rows, dim), dtype=tdt, device="cuda")
inv = torch.empty((rows,), dtype=torch.float32, device="cuda")
h = C.c_void_p()
vp, i64 = C.c_void_p, C.c_int64


def chk(rc):
    if rc != 0:
        raise RuntimeError(L.vm_last_error().decode())


st = torch.cuda.current_stream().cuda_stream
chk(L.vm_store_attach(C.byref(h), 0, dim, code, i64(rows), vp(X.data_ptr()), vp(inv.data_ptr())))
chk(L.vm_synth_fill(0, vp(X.data_ptr()), code, C.c_uint64(3), i64(0), i64(rows), dim, C.c_uint64(0), vp(st)))
chk(L.vm_store_set_size(h, i64(rows), i64(0), vp(st)))
g = torch.Generator().manual_seed(7)
q = (torch.randint(-127, 128, (nq, dim), generator=g).float() / 128.0).cuda()
oi = torch.empty((nq, k), dtype=torch.int64, device="cuda")
os_ = torch.empty((nq, k), dtype=torch.float64, device="cuda")
oc = torch.empty((nq,), dtype=torch.int32, device="cuda")


def call(fl):
    chk(L.vm_topk(h, vp(q.data_ptr()), 0, 1, nq, k, C.c_double(float("-inf")), 0, 1, fl, vp(oi.data_ptr()), vp(os_.data_ptr()),
                  vp(oc.data_ptr()), 1, None, vp(st)))


for _ in range(3):
    call(1 | flags)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    call(1 | flags)
e1.record()
torch.cuda.synchronize()
step = e0.elapsed_time(e1) / iters
scan = []
ms = C.c_float()
for _ in range(iters):
    call(1 | 16 | flags)
    chk(L.vm_store_last_scan_ms(h, C.byref(ms)))
    scan.append(ms.value)
scan.sort()
print(f"{lib_path.split('/')[-1]:28s} rows={rows} {dt} nq={nq} flags={flags}: scan median {scan[len(scan)//2]:.4f} ms (min {scan[0]:.4f}), step {step:.4f} ms, "
      f"{(rows * dim * (4 if dt == 'f32' else 2) + rows * 4) / scan[len(scan)//2] / 1e6:.0f} GB/s")
L.vm_store_destroy(h)
