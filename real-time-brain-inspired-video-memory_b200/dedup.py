"""All-pairs threshold scorer (entity / relation dedup) -- host wrapper of vm_pairs_above.

Replaces `Graph._are_same_context` (src/pipeline/prune.py:67-79): S = cosine_similarity(E);
fill_diagonal(S, 0); S > threshold -- generalised from `any()` to the set of pairs (i < j)."""
from __future__ import annotations

import ctypes as C
from typing import Tuple

import numpy as np
import torch

from . import _lib as L


def pairs_above(x: torch.Tensor, threshold: float, cap: int = 1 << 20, part: int = 0, nparts: int = 1
                ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """x: CUDA tensor [n, dim], float32 or bfloat16.  -> (i, j, score) numpy arrays sorted by (i, j);
    raises VidmemError(VM_ERR_OVERFLOW) if more than `cap` pairs exceed the threshold."""
    if not x.is_cuda or x.dim() != 2:
        raise TypeError("x must be a 2-D CUDA tensor")
    lib = L.load()
    n, dim = x.shape
    dt = {torch.float32: L.VM_F32, torch.bfloat16: L.VM_BF16}[x.dtype]
    ld = lib.vm_ld(dim)
    if ld != dim or not x.is_contiguous():
        xp = torch.zeros((n, ld), dtype=x.dtype, device=x.device)
        xp[:, :dim] = x
        x = xp
    dev = x.device
    oi = torch.empty((cap,), dtype=torch.int64, device=dev)
    oj = torch.empty((cap,), dtype=torch.int64, device=dev)
    os_ = torch.empty((cap,), dtype=torch.float32, device=dev)
    cnt = torch.zeros((1,), dtype=torch.int64, device=dev)
    L.check(lib.vm_pairs_above(dev.index or 0, x.data_ptr(), dt, n, dim, C.c_float(threshold), cap, oi.data_ptr(),
                               oj.data_ptr(), os_.data_ptr(), cnt.data_ptr(), part, nparts, 0,
                               torch.cuda.current_stream(dev).cuda_stream))
    total = int(cnt.item())
    if total > cap:
        raise L.VidmemError(L.VM_ERR_OVERFLOW, f"{total} pairs above the threshold exceed cap={cap}")
    i, j, s = oi[:total].cpu().numpy(), oj[:total].cpu().numpy(), os_[:total].cpu().numpy()
    order = np.lexsort((j, i))
    return i[order], j[order], s[order]


def _padded(x: torch.Tensor, lib) -> torch.Tensor:
    n, dim = x.shape
    ld = lib.vm_ld(dim)
    if ld != dim or not x.is_contiguous():
        xp = torch.zeros((n, ld), dtype=x.dtype, device=x.device)
        xp[:, :dim] = x
        return xp
    return x


def pairs_above_sharded(x: torch.Tensor, threshold: float, cap: int = 1 << 20, comm=None, root: int = -1):
    """Multi-GPU form: the upper-triangular tile grid is dealt cyclically to the ranks and every rank
    returns the full, (i, j)-sorted pair set.

    comm = a sharded.Communicator: ONE call into the C ABI (vm_pairs_above_sharded) -- operand replicated with an
    ncclBroadcast from `root` (root = -1: `x` is already the same on every rank), per-rank hit lists exchanged with
    ncclAllGather and concatenated on the device; nothing but the 8-byte counts touches the host.
    comm = None: the same exchange through torch.distributed (counts first, then a padded all_gather) -- the
    host-side restatement the gloo tests run on CPU."""
    if comm is not None:
        if not x.is_cuda or x.dim() != 2:
            raise TypeError("x must be a 2-D CUDA tensor")
        lib = L.load()
        n, dim = x.shape
        dt = {torch.float32: L.VM_F32, torch.bfloat16: L.VM_BF16}[x.dtype]
        xp = _padded(x, lib)
        dev = x.device
        oi = torch.empty((cap,), dtype=torch.int64, device=dev)
        oj = torch.empty((cap,), dtype=torch.int64, device=dev)
        os_ = torch.empty((cap,), dtype=torch.float32, device=dev)
        cnt = torch.zeros((1,), dtype=torch.int64, device=dev)
        L.check(lib.vm_pairs_above_sharded(comm.handle, xp.data_ptr(), dt, n, dim, C.c_float(threshold), cap, oi.data_ptr(),
                                           oj.data_ptr(), os_.data_ptr(), cnt.data_ptr(), int(root), 0,
                                           torch.cuda.current_stream(dev).cuda_stream))
        if root >= 0 and xp is not x:
            x.copy_(xp[:, :dim])                      # the broadcast landed in the padded copy
        total = int(cnt.item())
        i, j, s = oi[:total].cpu().numpy(), oj[:total].cpu().numpy(), os_[:total].cpu().numpy()
        order = np.lexsort((j, i))
        return i[order], j[order], s[order]
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    i, j, s = pairs_above(x, threshold, cap=cap, part=rank, nparts=world)
    dev = x.device
    cnt = torch.tensor([len(i)], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(counts, cnt)
    m = int(max(int(c.item()) for c in counts))
    buf = torch.zeros((m, 3), dtype=torch.float64, device=dev)
    if len(i):
        buf[:len(i), 0] = torch.from_numpy(i).to(dev)
        buf[:len(i), 1] = torch.from_numpy(j).to(dev)
        buf[:len(i), 2] = torch.from_numpy(s).to(dev)
    bufs = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(bufs, buf)
    parts = [b[:int(c.item())].cpu().numpy() for b, c in zip(bufs, counts)]
    allp = np.concatenate(parts) if parts else np.zeros((0, 3))
    gi, gj, gs = allp[:, 0].astype(np.int64), allp[:, 1].astype(np.int64), allp[:, 2].astype(np.float32)
    order = np.lexsort((gj, gi))
    return gi[order], gj[order], gs[order]
