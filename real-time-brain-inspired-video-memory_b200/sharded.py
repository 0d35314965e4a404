"""Row-sharded store over the GPUs of one NVSwitch box: one process per GPU, one shard each.

The only data-path collective is ONE ncclAllGather of the per-rank exact top-k lists
((score, global row)[nq][k], ~10 KB per rank) followed by a device-side merge inside
libvidmem (vm_topk_sharded).  torch.distributed is used for rendezvous only: it ships the
128-byte NCCL unique id from rank 0 to the other ranks.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Tuple

import numpy as np

from . import _lib as L


def shard_bounds(rows_total: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous row ranges: rank r owns [r*N/G, (r+1)*N/G)."""
    return [(rows_total * r // world, rows_total * (r + 1) // world) for r in range(world)]


def owner_of(row: int, rows_total: int, world: int) -> int:
    """Rank that owns a global row under shard_bounds()."""
    r = min(world - 1, (row * world) // max(rows_total, 1))
    while row < rows_total * r // world:
        r -= 1
    while row >= rows_total * (r + 1) // world:
        r += 1
    return r


def merge_lists_host(lists, k: int):
    """Host restatement of the device merge (tests of the N>1 logic on CPU/gloo): `lists` is a
    sequence of (idx [nq,k], score [nq,k], count [nq]) per rank, idx already global.
    Best k per query by (score desc, global row asc)."""
    nq = lists[0][0].shape[0]
    out_i = np.full((nq, k), -1, np.int64)
    out_s = np.zeros((nq, k), np.float64)
    out_c = np.zeros(nq, np.int32)
    for q in range(nq):
        rows, scores = [], []
        for idx, score, count in lists:
            c = int(count[q])
            rows.extend(idx[q, :c].tolist())
            scores.extend(score[q, :c].tolist())
        order = sorted(range(len(rows)), key=lambda e: (-scores[e], rows[e]))[:k]
        out_c[q] = len(order)
        for j, e in enumerate(order):
            out_i[q, j], out_s[q, j] = rows[e], scores[e]
    return out_i, out_s, out_c


class Communicator:
    """Owns a vm_comm (NCCL communicator) for this rank."""

    def __init__(self, device: int, nranks: int, rank: int, unique_id: bytes):
        self.lib = L.load()
        h = C.c_void_p()
        buf = C.create_string_buffer(unique_id, 128)
        L.check(self.lib.vm_comm_init_rank(C.byref(h), device, nranks, rank, buf))
        self.handle, self.nranks, self.rank = h, nranks, rank

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        L.check(L.load().vm_comm_unique_id(buf))
        return buf.raw

    @classmethod
    def from_torch_distributed(cls, device: int) -> "Communicator":
        import torch
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()
        uid = [cls.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        torch.cuda.set_device(device)
        return cls(device, world, rank, uid[0])

    def enable_peer_exchange(self) -> bool:
        """Switches the cross-shard exchange from ncclAllGather to the peer-memory kernel: allocates this
        rank's exchange buffer in torch symmetric memory (NVLink-mapped into every rank of the node),
        zero-fills it and hands the peer pointers to the library.  Returns False (NCCL path stays) when
        symmetric memory is unavailable."""
        import torch
        import torch.distributed as dist
        try:
            import torch.distributed._symmetric_memory as symm
            nbytes = int(self.lib.vm_comm_exchange_bytes())
            buf = symm.empty(nbytes, dtype=torch.uint8, device=torch.device("cuda", torch.cuda.current_device()))
            hdl = symm.rendezvous(buf, dist.group.WORLD)
            buf.zero_()
            torch.cuda.synchronize()
            dist.barrier()
            ptrs = (C.c_void_p * self.nranks)(*[int(p) for p in hdl.buffer_ptrs])
            L.check(self.lib.vm_comm_attach_peer_buffers(self.handle, ptrs, self.nranks))
            self._xchg = (buf, hdl)   # keep the mapping alive
            return True
        except Exception as e:  # pragma: no cover - depends on the platform
            self.peer_exchange_error = repr(e)
            return False

    def close(self):
        if self.handle:
            self.lib.vm_comm_destroy(self.handle)
            self.handle = None
