"""ctypes front-end of the C oracle (oracle/vm_oracle.c) plus the blocked large-N tier.

TEST INFRASTRUCTURE ONLY (see the header of vm_oracle.c): importable from tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libvm_oracle.so")
SUM_NAIVE, SUM_NEUMAIER = 0, 1

_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "vm_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        dp, ip, fp, bp = C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_float), C.POINTER(C.c_uint8)
        i64, dbl, cint = C.c_int64, C.c_double, C.c_int
        for name in ("vo_cosine_injector", "vo_cosine_retriever", "vo_cosine_utils"):
            f = getattr(L, name)
            f.restype = dbl
            f.argtypes = [dp, i64, dp, i64, cint]
        L.vo_batch_similarities.restype = cint
        L.vo_batch_similarities.argtypes = [dp, bp, i64, dp, bp, i64, i64, i64, cint, ip, dp, ip]
        L.vo_merge_max_by_id.restype = i64
        L.vo_merge_max_by_id.argtypes = [ip, dp, ip, i64, i64, i64, ip, dp]
        L.vo_vector_search.restype = cint
        L.vo_vector_search.argtypes = [dp, dp, bp, i64, i64, i64, dbl, cint, ip, dp, ip]
        L.vo_threshold_filter_ge.restype = i64
        L.vo_threshold_filter_ge.argtypes = [dp, dp, i64, i64, dbl, i64, cint, ip, dp]
        L.vo_pairs_above.restype = i64
        L.vo_pairs_above.argtypes = [fp, i64, i64, C.c_float, i64, ip, ip, fp]
        L.vo_normalize_rows_f32.restype = None
        L.vo_normalize_rows_f32.argtypes = [fp, i64, i64, fp]
        L.vo_pairs_rescore.restype = None
        L.vo_pairs_rescore.argtypes = [fp, i64, ip, ip, i64, fp]
        L.vo_representative.restype = i64
        L.vo_representative.argtypes = [fp, i64, i64, fp]
        L.vo_rescore_rows_f32.restype = None
        L.vo_rescore_rows_f32.argtypes = [fp, fp, i64, ip, i64, cint, dp]
        L.vo_synth_rows.restype = None
        L.vo_synth_rows.argtypes = [C.c_uint64, i64, i64, i64, C.c_uint64, fp]
        _lib = L
    return _lib


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def cosine(v1, v2, variant: str = "injector", sum_mode: int = SUM_NEUMAIER) -> float:
    a, b = _d(v1), _d(v2)
    f = getattr(lib(), "vo_cosine_" + variant)
    return float(f(_ptr(a, C.c_double), a.size, _ptr(b, C.c_double), b.size, sum_mode))


def batch_similarities(queries, store, top_k: int, query_ok=None, row_ok=None,
                       sum_mode: int = SUM_NEUMAIER) -> List[List[Tuple[int, float]]]:
    """Restates PreLLMInjector._calculate_batch_similarities (pre_llm_injector.py:346-372):
    per query, a list of (store_row, score) best-first, at most top_k long."""
    qs, st = _d(queries), _d(store)
    q, d = qs.shape if qs.ndim == 2 else (0, st.shape[1] if st.ndim == 2 else 0)
    n = st.shape[0] if st.ndim == 2 else 0
    oi = np.zeros((max(q, 1), max(top_k, 1)), np.int64)
    os_ = np.zeros((max(q, 1), max(top_k, 1)), np.float64)
    oc = np.zeros(max(q, 1), np.int64)
    qok = None if query_ok is None else np.ascontiguousarray(query_ok, np.uint8)
    rok = None if row_ok is None else np.ascontiguousarray(row_ok, np.uint8)
    rc = lib().vo_batch_similarities(_ptr(qs, C.c_double), _ptr(qok, C.c_uint8), q, _ptr(st, C.c_double),
                                     _ptr(rok, C.c_uint8), n, d, top_k, sum_mode, _ptr(oi, C.c_int64),
                                     _ptr(os_, C.c_double), _ptr(oc, C.c_int64))
    if rc != 0:
        raise MemoryError("vo_batch_similarities")
    return [[(int(oi[i, j]), float(os_[i, j])) for j in range(int(oc[i]))] for i in range(q)]


def merge_max_by_id(per_query: List[List[Tuple[int, float]]], top_k2: int) -> List[Tuple[int, float]]:
    """Restates the cross-query merge (pre_llm_injector.py:235-249)."""
    q = len(per_query)
    k = max([len(x) for x in per_query] + [1])
    idx = np.zeros((max(q, 1), k), np.int64)
    sc = np.zeros((max(q, 1), k), np.float64)
    cnt = np.zeros(max(q, 1), np.int64)
    for i, lst in enumerate(per_query):
        cnt[i] = len(lst)
        for j, (r, s) in enumerate(lst):
            idx[i, j], sc[i, j] = r, s
    oi = np.zeros(max(top_k2, 1), np.int64)
    os_ = np.zeros(max(top_k2, 1), np.float64)
    m = lib().vo_merge_max_by_id(_ptr(idx, C.c_int64), _ptr(sc, C.c_double), _ptr(cnt, C.c_int64), q, k, top_k2,
                                 _ptr(oi, C.c_int64), _ptr(os_, C.c_double))
    return [(int(oi[j]), float(os_[j])) for j in range(m)]


def vector_search(query, store, top_k: int, min_score: float = 0.3, row_ok=None,
                  sum_mode: int = SUM_NEUMAIER) -> List[Tuple[int, float]]:
    """Restates the Cypher of _vector_search_chunks (retriever_hybrid.py:293-306). PARITY UNPINNED."""
    qv, st = _d(query), _d(store)
    n, d = st.shape
    oi = np.zeros(max(top_k, 1), np.int64)
    os_ = np.zeros(max(top_k, 1), np.float64)
    oc = np.zeros(1, np.int64)
    rok = None if row_ok is None else np.ascontiguousarray(row_ok, np.uint8)
    lib().vo_vector_search(_ptr(qv, C.c_double), _ptr(st, C.c_double), _ptr(rok, C.c_uint8), n, d, top_k,
                           float(min_score), sum_mode, _ptr(oi, C.c_int64), _ptr(os_, C.c_double),
                           _ptr(oc, C.c_int64))
    return [(int(oi[j]), float(os_[j])) for j in range(int(oc[0]))]


def threshold_filter_ge(query, segs, threshold: float, top_k: int, sum_mode: int = SUM_NEUMAIER):
    """Restates the post-compression filter (retriever_hybrid.py:492-509)."""
    qv, sg = _d(query), _d(segs)
    n, d = sg.shape
    oi = np.zeros(max(top_k, 1), np.int64)
    os_ = np.zeros(max(top_k, 1), np.float64)
    m = lib().vo_threshold_filter_ge(_ptr(qv, C.c_double), _ptr(sg, C.c_double), n, d, float(threshold), top_k,
                                     sum_mode, _ptr(oi, C.c_int64), _ptr(os_, C.c_double))
    return [(int(oi[j]), float(os_[j])) for j in range(m)]


def pairs_above(x, threshold: float, cap: Optional[int] = None):
    """Restates prune.py:67-79 generalised to the pair set: (i, j, score) arrays, i < j, (i, j) order."""
    xf = np.ascontiguousarray(x, np.float32)
    n, d = xf.shape
    cap = int(cap if cap is not None else max(1024, 64 * n))
    oi = np.zeros(cap, np.int64)
    oj = np.zeros(cap, np.int64)
    os_ = np.zeros(cap, np.float32)
    total = lib().vo_pairs_above(_ptr(xf, C.c_float), n, d, np.float32(threshold), cap, _ptr(oi, C.c_int64),
                                 _ptr(oj, C.c_int64), _ptr(os_, C.c_float))
    if total > cap:
        return pairs_above(x, threshold, cap=int(total))
    return oi[:total].copy(), oj[:total].copy(), os_[:total].copy()


def pairs_above_streamed(x, threshold: float, block: int = 8192):
    """The all-pairs restatement (vo_pairs_above, prune.py:67-79) for row counts where the O(N^2 D) scalar loop takes
    hours (C4 at 262 144 rows and beyond).  A float32 BLAS product of the normalised rows pre-scores every block pair of
    the upper triangle (|error| <= PRE_TOL = (d + 8) 2^-24 in cosine units, Cauchy-Schwarz on unit rows); every pair whose
    pre-score exceeds threshold - 2 PRE_TOL is then decided by the oracle's own arithmetic (vo_pairs_rescore: binary64
    accumulation of the float32-normalised rows, rounded once to float32, strict >).  A pair below that margin cannot
    pass the oracle's comparison, so the returned set is the complete oracle pair set, in (i, j) order like pairs_above."""
    xf = np.ascontiguousarray(x, np.float32)
    n, d = xf.shape
    if n <= 1:
        return np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0, np.float32)
    L = lib()
    xn = np.empty_like(xf)
    L.vo_normalize_rows_f32(_ptr(xf, C.c_float), n, d, _ptr(xn, C.c_float))
    thr32 = float(np.float32(threshold))
    cut = np.float32(thr32 - 2.0 * (d + 8) * 2.0 ** -24 * 1.01)
    ci, cj = [], []
    buf = np.empty((min(block, n), min(block, n)), np.float32)  # one product buffer, reused (no page faults per block)
    for b0 in range(0, n, block):
        a = xn[b0:b0 + block]
        for c0 in range(b0, n, block):
            b = xn[c0:c0 + block]
            s = buf.reshape(-1)[:a.shape[0] * b.shape[0]].reshape(a.shape[0], b.shape[0])
            np.matmul(a, b.T, out=s)
            if c0 == b0:
                np.fill_diagonal(s, -2.0)                       # prune.py:77 fill_diagonal; i > j is dropped below
            rows = np.flatnonzero(s.max(axis=1) > cut)          # few rows per block hold a candidate at all
            if rows.size == 0:
                continue
            r, c = np.nonzero(s[rows] > cut)
            gi, gj = rows[r] + b0, c + c0
            keep = gi < gj
            ci.append(gi[keep].astype(np.int64))
            cj.append(gj[keep].astype(np.int64))
    ci = np.concatenate(ci) if ci else np.zeros(0, np.int64)
    cj = np.concatenate(cj) if cj else np.zeros(0, np.int64)
    sc = np.empty(len(ci), np.float32)
    if len(ci):
        L.vo_pairs_rescore(_ptr(xn, C.c_float), d, _ptr(ci, C.c_int64), _ptr(cj, C.c_int64), len(ci), _ptr(sc, C.c_float))
    hit = sc > np.float32(thr32)
    ci, cj, sc = ci[hit], cj[hit], sc[hit]
    order = np.lexsort((cj, ci))
    return ci[order], cj[order], sc[order]


def representative(x) -> Tuple[int, np.ndarray]:
    xf = np.ascontiguousarray(x, np.float32)
    n, d = xf.shape
    sims = np.zeros(n, np.float32)
    i = lib().vo_representative(_ptr(xf, C.c_float), n, d, _ptr(sims, C.c_float))
    return int(i), sims


def synth_rows_c(seed: int, row0: int, n: int, d: int, dup_period: int = 0) -> np.ndarray:
    out = np.empty((n, d), np.float32)
    lib().vo_synth_rows(seed, row0, n, d, dup_period, _ptr(out, C.c_float))
    return out


def rescore_rows(query_f32, store_f32, rows, sum_mode: int = SUM_NEUMAIER) -> np.ndarray:
    qf = np.ascontiguousarray(query_f32, np.float32)
    st = np.ascontiguousarray(store_f32, np.float32)
    r = np.ascontiguousarray(rows, np.int64)
    out = np.empty(len(r), np.float64)
    lib().vo_rescore_rows_f32(_ptr(qf, C.c_float), _ptr(st, C.c_float), st.shape[1], _ptr(r, C.c_int64), len(r),
                              sum_mode, _ptr(out, C.c_double))
    return out


def topk_blocked(queries_f32, store_f32, top_k: int, slack: int = 24, block: int = 262144,
                 sum_mode: int = SUM_NEUMAIER):
    """Large-N tier (SURVEY.md 8c tier-2): a float64 BLAS matmul per row block pre-ranks
    (|error| ~ 1e-15, far below any gap the final order depends on unless more than `slack`
    rows tie at the boundary), then the top (top_k + slack) rows per query are rescored with
    the bit-exact reference formula and ordered by (score desc, row asc).
    Returns (idx [q, top_k] int64, score [q, top_k] float64, count [q])."""
    qf = np.ascontiguousarray(queries_f32, np.float32)
    st = np.ascontiguousarray(store_f32, np.float32)
    q, d = qf.shape
    n = st.shape[0]
    keep = min(n, top_k + slack)
    qd = qf.astype(np.float64)
    qn = np.sqrt((qd * qd).sum(1))
    qn[qn == 0] = 1.0
    cand_s = np.full((q, 0), -np.inf)
    cand_i = np.zeros((q, 0), np.int64)
    for b0 in range(0, n, block):
        xb = st[b0:b0 + block].astype(np.float64)
        nb = np.sqrt((xb * xb).sum(1))
        nb[nb == 0] = 1.0
        s = (qd @ xb.T) / qn[:, None] / nb[None, :]
        cs = np.concatenate([cand_s, s], axis=1)
        ci = np.concatenate([cand_i, np.broadcast_to(np.arange(b0, b0 + xb.shape[0]), s.shape)], axis=1)
        if cs.shape[1] > keep:
            part = np.argpartition(-cs, keep - 1, axis=1)[:, :keep]
            cs = np.take_along_axis(cs, part, 1)
            ci = np.take_along_axis(ci, part, 1)
        cand_s, cand_i = cs, ci
    oi = np.zeros((q, top_k), np.int64)
    os_ = np.zeros((q, top_k), np.float64)
    oc = np.zeros(q, np.int64)
    for i in range(q):
        rows = np.sort(cand_i[i])
        ex = rescore_rows(qf[i], st, rows, sum_mode)
        order = np.lexsort((rows, -ex))
        m = min(top_k, len(rows))
        oi[i, :m] = rows[order[:m]]
        os_[i, :m] = ex[order[:m]]
        oc[i] = m
    return oi, os_, oc


def topk_streamed(queries_f32, n: int, block_fn, rows_fn, top_k: int, slack: int = 54, block: int = 262144,
                  sum_mode: int = SUM_NEUMAIER):
    """The blocked tier for stores too large to hold on the host as one array (C3's 12.5 M-row shard and beyond).
    block_fn(b0, b1) -> float32 rows [b0, b1); rows_fn(sorted row indices) -> float32 rows.  Pre-ranking is a float32
    BLAS matmul per block (|error| <= PRE_TOL = (d + 8) 2^-24 in cosine units), the best top_k + slack rows per
    query are rescored with the bit-exact reference formula, and the result is only returned if it is PROVEN complete:
    every row that was not rescored has a pre-ranking score more than 2 * PRE_TOL below the k-th exact score (else
    AssertionError -- raise `slack`).  Returns (idx [q, top_k] int64, score [q, top_k] float64, count [q])."""
    qf = np.ascontiguousarray(queries_f32, np.float32)
    q, d = qf.shape
    PRE_TOL = (d + 8) * 2.0 ** -24 * 1.01      # float32 dot product + normalisations, Cauchy-Schwarz, cosine units
    keep = min(n, top_k + slack)
    qn = np.sqrt(np.einsum("ij,ij->i", qf, qf, dtype=np.float64))
    qn[qn == 0] = 1.0
    qs = (qf / qn[:, None]).astype(np.float32)
    cand_s = np.full((q, 0), -np.inf, np.float32)
    cand_i = np.zeros((q, 0), np.int64)
    for b0 in range(0, n, block):
        xb = np.ascontiguousarray(block_fn(b0, min(n, b0 + block)), np.float32)
        nb = np.sqrt(np.einsum("ij,ij->i", xb, xb, dtype=np.float64))
        nb[nb == 0] = 1.0
        s = (qs @ xb.T) / nb[None, :].astype(np.float32)
        cs = np.concatenate([cand_s, s], axis=1)
        ci = np.concatenate([cand_i, np.broadcast_to(np.arange(b0, b0 + xb.shape[0]), s.shape)], axis=1)
        if cs.shape[1] > keep:
            part = np.argpartition(-cs, keep - 1, axis=1)[:, :keep]
            cs = np.take_along_axis(cs, part, 1)
            ci = np.take_along_axis(ci, part, 1)
        cand_s, cand_i = cs, ci
    oi = np.zeros((q, top_k), np.int64)
    os_ = np.zeros((q, top_k), np.float64)
    oc = np.zeros(q, np.int64)
    for i in range(q):
        rows = np.sort(cand_i[i])
        vals = np.ascontiguousarray(rows_fn(rows), np.float32)
        ex = rescore_rows(qf[i], vals, np.arange(len(rows)), sum_mode)
        order = np.lexsort((rows, -ex))
        m = min(top_k, len(rows))
        oi[i, :m] = rows[order[:m]]
        os_[i, :m] = ex[order[:m]]
        oc[i] = m
        if len(rows) < n and m == top_k:      # rows outside the candidate set: pre-ranking score <= the set's minimum
            assert float(cand_s[i].min()) + 2 * PRE_TOL < os_[i, m - 1], "oracle.topk_streamed: candidate set not provably complete"
    return oi, os_, oc

