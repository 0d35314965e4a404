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


class ShardedChunkStore:
    """Streaming inserts over the GPUs of one box ("streaming inserts"): every rank runs
    the same call sequence (SPMD); each insert batch's NEW ids go to the least-full rank as one
    block, an id that is already known is overwritten on the rank that owns it (the reference
    MERGEs by id, src/components/neo4j_handler.py:229).  No collective on the insert path.

    Every rank keeps the whole id table (global row -> chunk id, owner rank); only the owner holds
    the embedding in HBM.  Global row == first-seen order over all ranks == the reference's dict
    insertion order, so the tie rule "earliest row wins"  holds across shards: local
    rows are in increasing global order, and the cross-rank merge orders by (score desc, global
    row asc).

    Queries: local exact top-k (vm_topk) -> local rows renamed to global rows -> ONE all-gather of
    the [nq][k] lists -> device merge (vm_merge_topk_lists); identical results on every rank.
    Exposes the ResidentChunkStore surface the adapters use (upsert / sync_from_dict / topk / ids /
    row_of / meta / device), so it plugs into ChunkSimilarityBackend and VectorSearchBackend."""

    def __init__(self, dtype: str = "f64", device: int = 0, rank: int = None, world: int = None, group=None,
                 initial_capacity: int = 8192):
        import torch.distributed as dist
        from .adapters import ResidentChunkStore
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else int(rank)
        self.world = dist.get_world_size(group) if world is None else int(world)
        self.device = device
        self.local = ResidentChunkStore(dtype, device, initial_capacity)
        self.ids: List[str] = []            # global row -> chunk id
        self.row_of = {}                    # chunk id -> global row
        self.owner: List[int] = []          # global row -> rank
        self.load = [0] * self.world        # rows per rank
        self.meta = self.local.meta         # replicated on every rank (host only)
        self._mirror_prev = None            # last dict handed to sync_from_dict

    def __len__(self) -> int:
        return len(self.ids)

    def clear(self) -> None:
        self.local.clear()
        self.ids, self.row_of, self.owner = [], {}, []
        self.load = [0] * self.world
        self._mirror_prev = None

    def upsert(self, items, meta=None) -> None:
        merged = {}
        for cid, emb in items:              # an id written twice in one batch is one row: first position, last value
            merged[cid] = emb
        items = list(merged.items())
        fresh = [cid for cid, _ in items if cid not in self.row_of]
        target = min(range(self.world), key=lambda r: (self.load[r], r)) if fresh else -1
        for cid in fresh:
            self.row_of[cid] = len(self.ids)
            self.ids.append(cid)
            self.owner.append(target)
        if fresh:
            self.load[target] += len(fresh)
        mine = [(cid, emb) for cid, emb in items if self.owner[self.row_of[cid]] == self.rank]
        if mine:
            self.local.upsert(mine)
        if meta:
            self.meta.update(meta)

    def sync_from_dict(self, existing) -> None:
        """Mirror mode, same contract as ResidentChunkStore.sync_from_dict (changed values are rewritten on
        their owner, new keys are routed, a removed or reordered key rebuilds the store)."""
        keys = list(existing.keys())
        n = len(self.ids)
        prev = self._mirror_prev
        if keys[:n] != self.ids:
            self.clear()
            n, prev = 0, None
        changed = []
        if n:
            changed = keys[:n] if prev is None else [c for c in keys[:n] if existing[c] is not prev.get(c) and existing[c] != prev.get(c)]
        if changed or len(keys) > n:
            self.upsert([(c, existing[c]) for c in changed] + [(c, existing[c]) for c in keys[n:]])
        self._mirror_prev = existing

    # -- the exchange -------------------------------------------------------------------------
    def _gather(self, idx: np.ndarray, score: np.ndarray, count: np.ndarray):
        """One all-gather of this rank's lists; -> per-rank (idx, score, count), idx global."""
        import torch
        import torch.distributed as dist
        nq, k = idx.shape
        packed = np.concatenate([idx.reshape(-1), score.reshape(-1).view(np.int64), count.astype(np.int64)])
        on_gpu = dist.get_backend(self.group) == "nccl"
        t = torch.from_numpy(packed)
        if on_gpu:
            t = t.to(torch.device("cuda", self.device))
        out = torch.empty((self.world * packed.size,), dtype=torch.int64, device=t.device)
        dist.all_gather_into_tensor(out, t, group=self.group)
        return out.view(self.world, packed.size), nq, k

    def _merge(self, gathered, nq: int, k: int):
        """Device merge of the gathered lists (vm_merge_topk_lists): (score desc, global row asc)."""
        import torch
        lib = L.load()
        dev = torch.device("cuda", self.device)
        g = gathered.to(dev)
        idx = g[:, :nq * k].contiguous()
        score = g[:, nq * k:2 * nq * k].contiguous().view(torch.float64)
        count = g[:, 2 * nq * k:].to(torch.int32).contiguous()
        o_idx = torch.empty((nq, k), dtype=torch.int64, device=dev)
        o_score = torch.empty((nq, k), dtype=torch.float64, device=dev)
        o_count = torch.empty((nq,), dtype=torch.int32, device=dev)
        L.check(lib.vm_merge_topk_lists(self.device, idx.data_ptr(), score.data_ptr(), count.data_ptr(), self.world, nq, k,
                                        o_idx.data_ptr(), o_score.data_ptr(), o_count.data_ptr(),
                                        torch.cuda.current_stream(dev).cuda_stream))
        return o_idx.cpu().numpy(), o_score.cpu().numpy(), o_count.cpu().numpy()

    def _local_lists(self, queries, k: int, min_score: float, score_mode: int, flags: int):
        """This rank's exact top-k per query with rows renamed to GLOBAL rows: (idx [nq,k], score, count)."""
        nq = len(queries)
        local = self.local.topk(queries, k, min_score=min_score, score_mode=score_mode, flags=flags)
        idx = np.full((nq, k), -1, np.int64)
        score = np.zeros((nq, k), np.float64)
        count = np.zeros(nq, np.int32)
        for i, lst in enumerate(local):
            count[i] = len(lst)
            for j, (cid, s) in enumerate(lst):
                idx[i, j], score[i, j] = self.row_of[cid], s
        return idx, score, count

    def _named(self, m_idx, m_score, m_count):
        return [[(self.ids[int(m_idx[i, j])], float(m_score[i, j])) for j in range(int(m_count[i]))] for i in range(len(m_count))]

    def topk(self, queries, k: int, min_score: float = -np.inf, score_mode: int = L.VM_SCORE_RAW, flags: int = 0):
        """Same contract as ResidentChunkStore.topk: list (per query) of [(chunk_id, score)]."""
        queries = list(queries)
        if len(queries) == 0 or not self.ids:
            return [[] for _ in queries]
        idx, score, count = self._local_lists(queries, k, min_score, score_mode, flags)
        gathered, nq, k = self._gather(idx, score, count)
        return self._named(*self._merge(gathered, nq, k))
