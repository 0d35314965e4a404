"""Test doubles shared by the CPU suites (never imported by the product)."""
import numpy as np

from oracle import oracle


class OracleBackedStore:
    """Stands in for vidmem_b200.store.EmbeddingStore inside ResidentChunkStore (tests only)."""

    exact = False
    growable = True

    def __init__(self, dim, capacity, dtype="f32", device=0, max_capacity=None):
        self.dim, self.capacity = dim, capacity
        self.max_capacity = max_capacity if max_capacity is not None else capacity
        self.dtype = dtype
        self.X = np.zeros((0, dim))
        self.ok = np.zeros(0, np.uint8)

    def __len__(self):
        return len(self.X)

    def _stored(self, rows):
        """what the store holds: a binary64 store ("f64...") keeps the values as given, the others round once"""
        rows = np.asarray(rows, np.float64)
        return rows if str(self.dtype).startswith("f64") else rows.astype(np.float32).astype(np.float64)

    def reserve(self, capacity):
        assert capacity <= self.max_capacity
        self.capacity = max(self.capacity, capacity)

    # what ResidentChunkStore reads directly from the real store (torch tensors there)
    @property
    def inv_norms(self):
        import torch
        return torch.from_numpy(np.where(self.ok > 0, 1.0, -1.0).astype(np.float32))

    @property
    def rows(self):
        import torch
        return torch.from_numpy(self.X.astype(np.float32))

    def append(self, rows):
        first = len(self.X)
        assert first + len(rows) <= self.capacity, "append beyond the backed capacity"
        rows = self._stored(rows)
        self.X = np.concatenate([self.X, rows]); self.ok = np.concatenate([self.ok, np.ones(len(rows), np.uint8)])
        return first

    def update(self, row0, rows):
        self.X[row0:row0 + len(rows)] = self._stored(rows); self.ok[row0:row0 + len(rows)] = 1

    def invalidate(self, rows):
        self.ok[list(rows)] = 0

    def clear(self):
        self.X = self.X[:0]; self.ok = self.ok[:0]

    def topk(self, q, k, min_score=-np.inf, score_mode=0, flags=0):
        if score_mode == 1:                                            # VM_SCORE_NEO4J: (1 + cos) / 2, strict > min_score
            res = [oracle.vector_search(qi, self.X, k, min_score=min_score, row_ok=self.ok) for qi in np.asarray(q, np.float64)]
        else:
            res = oracle.batch_similarities(q, self.X, k, row_ok=self.ok)
            if min_score > -np.inf:
                res = [[(r, s) for r, s in lst if s > min_score] for lst in res]
        idx = np.full((len(q), k), -1, np.int64); sc = np.zeros((len(q), k)); cnt = np.zeros(len(q), np.int32)
        for i, lst in enumerate(res):
            cnt[i] = len(lst)
            for j, (r, s) in enumerate(lst):
                idx[i, j], sc[i, j] = r, s
        return idx, sc, cnt

    def close(self):
        pass
