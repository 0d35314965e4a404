"""EmbeddingStore: one HBM-resident shard of chunk embeddings + the exact top-k scorer.

Host-side mirror of the reference's store for this path: where the reference re-fetches
`{chunk_id: embedding}` from Neo4j on every batch (src/components/pre_llm_injector.py:390-412)
and loops over it in Python (:356-370), this class keeps the rows resident on the GPU
(PyTorch owns the device memory and the stream; libvidmem.so does the arithmetic).
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib as L

_DT = {"f32": L.VM_F32, "fp32": L.VM_F32, "float32": L.VM_F32, "bf16": L.VM_BF16, "bfloat16": L.VM_BF16}
#: binary64 stores: the rows are kept as given (what the reference scores: Python floats), the scan reads a rounded
#: shadow copy -- "f64" with an fp32 shadow, "f64+bf16" with a bf16 shadow (half the scan traffic, a wider near-tie band)
_DT_EXACT = {"f64": L.VM_F32, "fp64": L.VM_F32, "float64": L.VM_F32, "f64+f32": L.VM_F32, "f64+bf16": L.VM_BF16}
_TORCH_DT = {L.VM_F32: torch.float32, L.VM_BF16: torch.bfloat16}


def _np_dtype_code(a: np.ndarray) -> int:
    if a.dtype == np.float32:
        return L.VM_F32
    if a.dtype == np.float64:
        return L.VM_F64
    raise TypeError(f"unsupported host dtype {a.dtype}")


def _torch_dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return L.VM_F32
    if t.dtype == torch.bfloat16:
        return L.VM_BF16
    if t.dtype == torch.float64:
        return L.VM_F64
    raise TypeError(f"unsupported tensor dtype {t.dtype}")


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream_ptr(device: torch.device) -> int:
    # the caller's current stream on `device`; the raw accessor skips building a Stream object (~5 us per call)
    if _raw_stream is not None:
        return _raw_stream(device.index)
    return torch.cuda.current_stream(device).cuda_stream


def device_row_limit(dim: int, dtype: str, device: int = 0) -> int:
    """Rows of this shape the device's memory could hold at most: what a growable store reserves VIRTUAL addresses for
    when the caller names no maximum (reserving costs no memory)."""
    per_row = ((int(dim) + 7) & ~7) * {"f32": 4, "bf16": 2, "f64+bf16": 10}.get(dtype, 12) + 4
    total = torch.cuda.get_device_properties(device).total_memory if torch.cuda.is_available() else 1 << 34
    return max(1, min(total // per_row, (1 << 31) - 512))


class _DeviceRange:
    """A library-owned device range as a __cuda_array_interface__ object (torch.as_tensor wraps it without a copy)."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 3, "strides": None}


class EmbeddingStore:
    """Row-major embedding shard on one GPU.  Row index == append order (the reference's dict
    insertion order).

    max_capacity=None: a store over fixed torch buffers of `capacity` rows.  max_capacity=N: a GROWABLE store -- the
    library reserves virtual addresses for N rows and backs them with HBM as rows arrive (`append` grows on demand,
    `reserve` explicitly); resident rows are never copied and `rows` / `inv_norms` keep their addresses."""

    def __init__(self, dim: int, capacity: int, dtype: str = "f32", device: int = 0, _buffers=None,
                 max_capacity: Optional[int] = None):
        if not torch.cuda.is_available():
            raise RuntimeError("EmbeddingStore needs a CUDA device (sm_100); there is no CPU fallback")
        self.lib = L.load()
        self.dim, self.capacity = int(dim), int(capacity)
        self.growable = max_capacity is not None
        self.dtype = str(dtype)
        self.exact = dtype in _DT_EXACT
        self.dtype_code = _DT_EXACT[dtype] if self.exact else _DT[dtype]   # what the scan reads (`rows`)
        self.device = torch.device("cuda", device)
        self.ld = self.lib.vm_ld(self.dim)
        # PyTorch owns the HBM; the library attaches to it
        self.rows_exact = None
        if self.growable:
            h = C.c_void_p()
            with torch.cuda.device(self.device):
                L.check(self.lib.vm_store_create_growable(C.byref(h), device, self.dim, self.dtype_code, 1 if self.exact else 0,
                                                          self.capacity, max(int(max_capacity), self.capacity)))
            self._h = h
            self.max_capacity = int(self.lib.vm_store_max_capacity(h))
            self._refresh_views()
            self.last_stats = L.TopkStats()
            self._stats_ref = C.byref(self.last_stats)
            return
        self.max_capacity = self.capacity
        if _buffers is None:
            self.rows = torch.empty((self.capacity, self.ld), dtype=_TORCH_DT[self.dtype_code], device=self.device)
            self.inv_norms = torch.empty((self.capacity,), dtype=torch.float32, device=self.device)
            if self.exact:
                self.rows_exact = torch.empty((self.capacity, self.ld), dtype=torch.float64, device=self.device)
        elif self.exact:
            self.rows, self.inv_norms, self.rows_exact = _buffers
        else:
            self.rows, self.inv_norms = _buffers
        h = C.c_void_p()
        if self.exact:
            L.check(self.lib.vm_store_attach_exact(C.byref(h), device, self.dim, self.dtype_code, self.capacity,
                                                   self.rows.data_ptr(), self.inv_norms.data_ptr(), self.rows_exact.data_ptr()))
        else:
            L.check(self.lib.vm_store_attach(C.byref(h), device, self.dim, self.dtype_code, self.capacity,
                                             self.rows.data_ptr(), self.inv_norms.data_ptr()))
        self._h = h
        self.last_stats = L.TopkStats()
        self._stats_ref = C.byref(self.last_stats)

    def _refresh_views(self) -> None:
        """(growable store) torch views of the library-owned ranges, over the rows backed right now."""
        self.capacity = cap = int(self.lib.vm_store_capacity(self._h))
        dev = self.device
        if self.dtype_code == L.VM_BF16:
            r = torch.as_tensor(_DeviceRange(self.lib.vm_store_rows_ptr(self._h), (cap, self.ld), "<i2"), device=dev).view(torch.bfloat16)
        else:
            r = torch.as_tensor(_DeviceRange(self.lib.vm_store_rows_ptr(self._h), (cap, self.ld), "<f4"), device=dev)
        self.rows = r
        self.inv_norms = torch.as_tensor(_DeviceRange(self.lib.vm_store_inv_norms_ptr(self._h), (cap,), "<f4"), device=dev)
        if self.exact:
            self.rows_exact = torch.as_tensor(_DeviceRange(self.lib.vm_store_rows_exact_ptr(self._h), (cap, self.ld), "<f8"), device=dev)

    def reserve(self, capacity: int) -> None:
        """Back at least `capacity` rows (growable stores; no-op when already backed).  Nothing is copied."""
        if int(capacity) <= self.capacity:
            return
        if not self.growable:
            raise ValueError(f"store over fixed buffers of {self.capacity} rows cannot grow; create it with max_capacity")
        L.check(self.lib.vm_store_reserve(self._h, int(capacity)))
        self._refresh_views()

    def resident_bytes(self) -> int:
        """Physical HBM behind the store's rows, inverse norms and (binary64 store) original rows."""
        return int(self.lib.vm_store_resident_bytes(self._h))

    def prefix_view(self, n: int) -> "EmbeddingStore":
        """A second handle over the FIRST n resident rows (same HBM, same cached inverse norms): what a store
        holding only those rows would answer.  Read-only use; close it before the parent."""
        n = int(n)
        if not 1 <= n <= len(self):
            raise ValueError(f"prefix of {n} rows outside [1, {len(self)}]")
        bufs = (self.rows[:n], self.inv_norms[:n]) + ((self.rows_exact[:n],) if self.exact else ())
        v = EmbeddingStore(self.dim, n, self.dtype, self.device.index, _buffers=bufs)
        L.check(self.lib.vm_store_set_size(v._h, n, n, _stream_ptr(self.device)))   # norms are already cached
        return v

    # -- lifetime ---------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None):
            if self.growable:
                self.rows = self.inv_norms = self.rows_exact = None   # views of memory the library is about to unmap
            self.lib.vm_store_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self) -> int:
        return int(self.lib.vm_store_size(self._h))

    # -- writes -----------------------------------------------------------------------------
    def _src(self, rows) -> Tuple[int, int, int, int, object]:
        """-> (ptr, dtype_code, mem, n, keepalive)"""
        if isinstance(rows, torch.Tensor):
            t = rows.contiguous()
            if t.dim() != 2 or t.shape[1] != self.dim:
                raise ValueError(f"expected [n, {self.dim}] rows, got {tuple(t.shape)}")
            if t.is_cuda:
                return t.data_ptr(), _torch_dtype_code(t), L.VM_MEM_DEVICE, t.shape[0], t
            if t.dtype == torch.bfloat16:
                return t.data_ptr(), L.VM_BF16, L.VM_MEM_HOST, t.shape[0], t
            rows = t.numpy()
        a = np.asarray(rows)
        if a.dtype not in (np.float32, np.float64):
            a = a.astype(np.float64)
        a = np.ascontiguousarray(a)
        if a.ndim != 2 or a.shape[1] != self.dim:
            raise ValueError(f"expected [n, {self.dim}] rows, got {a.shape}")
        return a.ctypes.data, _np_dtype_code(a), L.VM_MEM_HOST, a.shape[0], a

    def append(self, rows) -> int:
        """Appends rows, returns the index of the first one (insert hook of
        src/components/neo4j_handler.py:221-253)."""
        ptr, dt, mem, n, keep = self._src(rows)
        first = C.c_int64(-1)
        L.check(self.lib.vm_store_append(self._h, ptr, dt, mem, n, C.byref(first), _stream_ptr(self.device)))
        if self.growable and int(first.value) + n > self.capacity:
            self._refresh_views()                                 # the library backed more rows
        if mem == L.VM_MEM_HOST:
            torch.cuda.current_stream(self.device).synchronize()  # host buffer may be released after return
        return int(first.value)

    def update(self, row0: int, rows) -> None:
        ptr, dt, mem, n, keep = self._src(rows)
        L.check(self.lib.vm_store_update(self._h, int(row0), ptr, dt, mem, n, _stream_ptr(self.device)))
        if mem == L.VM_MEM_HOST:
            torch.cuda.current_stream(self.device).synchronize()

    def invalidate(self, rows: Sequence[int]) -> None:
        a = np.ascontiguousarray(np.asarray(list(rows), dtype=np.int64))
        L.check(self.lib.vm_store_invalidate(self._h, a.ctypes.data, a.size, _stream_ptr(self.device)))
        torch.cuda.current_stream(self.device).synchronize()

    def set_size(self, n: int, recompute_from_row: int = 0) -> None:
        """Declare rows [0, n) of `self.rows` filled in place (e.g. by synth_fill)."""
        L.check(self.lib.vm_store_set_size(self._h, int(n), int(recompute_from_row), _stream_ptr(self.device)))

    def last_scan_ms(self) -> float:
        """Device time of the scan kernel(s) of the last top-k call made with VM_FLAG_TIMING."""
        ms = C.c_float(0.0)
        L.check(self.lib.vm_store_last_scan_ms(self._h, C.byref(ms)))
        return float(ms.value)

    def avg_scan_ms(self):
        """-> (mean scan-kernel ms, calls) over the VM_FLAG_TIMING calls since the previous read (at most the
        last 64): every call carries its own event pair, so a back-to-back loop is measured as it ran."""
        ms, n = C.c_float(0.0), C.c_int(0)
        L.check(self.lib.vm_store_avg_scan_ms(self._h, C.byref(ms), C.byref(n)))
        return float(ms.value), int(n.value)

    def counters(self, reset: bool = False) -> dict:
        """Certification counters since creation / the last reset (also cover VM_FLAG_ASYNC calls): batches, queries,
        uncertified, band_settled, collect_settled, full_rescans.  Synchronises with the device."""
        c = L.StoreCounters()
        L.check(self.lib.vm_store_read_counters(self._h, C.byref(c), 1 if reset else 0))
        return c.as_dict()

    def band_keys(self):
        """-> (kept [nq], spilled [nq]) int64 arrays: keys the last tcgen05 scan kept per query / spilled per query."""
        kept, spilled, n = (C.c_int64 * 64)(), (C.c_int64 * 64)(), C.c_int(0)
        L.check(self.lib.vm_store_band_keys(self._h, 64, kept, spilled, C.byref(n)))
        return np.array(kept[:n.value], np.int64), np.array(spilled[:n.value], np.int64)

    def clear(self) -> None:
        L.check(self.lib.vm_store_clear(self._h))

    def synth_fill(self, seed: int, n: int, row0: int = 0, dup_period: int = 0, first_buffer_row: int = 0) -> None:
        """Generates synthetic rows [row0, row0+n) of the SURVEY.md 8d generator directly in HBM at
        buffer rows [first_buffer_row, ...)."""
        dst = self.rows[first_buffer_row:first_buffer_row + n]
        L.check(self.lib.vm_synth_fill(self.device.index, dst.data_ptr(), self.dtype_code, seed, row0, n, self.dim,
                                       dup_period, _stream_ptr(self.device)))

    # -- persistence (binary sidecar instead of JSON float lists) ---------------------
    def save(self, path: str, ids: Optional[Sequence[str]] = None, extra: Optional[dict] = None) -> None:
        """Writes the resident rows as a binary sidecar: `<path>.npz` with the raw row values in the
        store dtype (bf16 as uint16 bit patterns), the skipped-row mask and, optionally, the chunk ids
        and caller strings (`extra`: name -> str).  The JSON export of
        src/components/graph_exporter.py:81-108 stores embeddings as float lists; this is the same
        information without float parsing.  Plain arrays only -- nothing in the file is pickled."""
        n = len(self)
        if self.exact:      # binary64 originals; the shadow is re-derived on load
            raw = self.rows_exact[:n, :self.dim].contiguous().cpu().numpy()
        else:
            rows = self.rows[:n, :self.dim].contiguous()
            raw = rows.view(torch.int16).cpu().numpy().view(np.uint16) if self.dtype_code == L.VM_BF16 else rows.cpu().numpy()
        skipped = (self.inv_norms[:n] < 0).cpu().numpy()
        arrays = {"rows": raw, "skipped": skipped, "dim": np.int64(self.dim), "dtype": np.int64(self.dtype_code),
                  "exact": np.int64(1 if self.exact else 0),
                  "ids": np.asarray([str(x) for x in ids] if ids is not None else [], dtype=np.str_)}
        for name, text in (extra or {}).items():
            arrays["extra_" + name] = np.frombuffer(str(text).encode("utf-8"), dtype=np.uint8)
        np.savez(path, **arrays)

    @classmethod
    def load(cls, path: str, capacity: Optional[int] = None, device: int = 0, with_extra: bool = False,
             min_capacity: int = 1, max_capacity: Optional[int] = None):
        """-> (store, ids) or (store, ids, extra).  Bit-identical rows, same skipped rows, same order."""
        z = np.load(path if path.endswith(".npz") else path + ".npz", allow_pickle=False)
        dim, code = int(z["dim"]), int(z["dtype"])
        raw = z["rows"]
        n = raw.shape[0]
        exact = "exact" in z.files and int(z["exact"]) == 1
        name = ("f64+bf16" if code == L.VM_BF16 else "f64") if exact else ("bf16" if code == L.VM_BF16 else "f32")
        if max_capacity == "auto":       # growable, bounded by what the device could hold
            max_capacity = device_row_limit(dim, name, device)
        cap = max(int(capacity or n), int(min_capacity), 1)
        st = cls(dim, cap, name, device, max_capacity=None if max_capacity is None else max(int(max_capacity), cap))
        if n:
            if code == L.VM_BF16 and not exact:
                t = torch.from_numpy(raw.view(np.int16).copy()).view(torch.bfloat16)
                st.append(t.to(st.device))
            else:
                st.append(raw)
            bad = np.nonzero(z["skipped"])[0]
            if len(bad):
                st.invalidate(bad)
        ids = [str(x) for x in z["ids"].tolist()]
        if not with_extra:
            return st, ids
        extra = {k[len("extra_"):]: bytes(z[k].tobytes()).decode("utf-8") for k in z.files if k.startswith("extra_")}
        return st, ids, extra

    # -- reads ------------------------------------------------------------------------------
    def topk(self, queries, k: int, min_score: float = -math.inf, score_mode: int = L.VM_SCORE_RAW,
             sum_mode: Optional[int] = None, flags: int = 0, comm=None, row_offset: int = 0):
        """Host in / host out.  -> (idx [nq,k] int64, score [nq,k] float64, count [nq] int32).
        Scores are bit-identical to the reference formula; ties -> lowest row."""
        if isinstance(queries, torch.Tensor) and queries.is_cuda:
            raise TypeError("use topk_device for CUDA tensors")
        q = np.asarray(queries.numpy() if isinstance(queries, torch.Tensor) else queries)
        if q.dtype not in (np.float32, np.float64):
            q = q.astype(np.float64)
        q = np.ascontiguousarray(q)
        if q.ndim != 2 or q.shape[1] != self.dim:
            raise ValueError(f"expected [nq, {self.dim}] queries, got {q.shape}")
        nq = q.shape[0]
        idx = np.full((nq, k), -1, np.int64)
        score = np.zeros((nq, k), np.float64)
        count = np.zeros((nq,), np.int32)
        sm = L.DEFAULT_SUM_MODE if sum_mode is None else sum_mode
        st = _stream_ptr(self.device)
        if comm is None:
            rc = self.lib.vm_topk(self._h, q.ctypes.data, _np_dtype_code(q), L.VM_MEM_HOST, nq, k, float(min_score),
                                  score_mode, sm, flags, idx.ctypes.data, score.ctypes.data, count.ctypes.data,
                                  L.VM_MEM_HOST, self._stats_ref, st)
        else:
            rc = self.lib.vm_topk_sharded(self._h, comm.handle, int(row_offset), q.ctypes.data, _np_dtype_code(q),
                                          L.VM_MEM_HOST, nq, k, float(min_score), score_mode, sm, flags,
                                          idx.ctypes.data, score.ctypes.data, count.ctypes.data, L.VM_MEM_HOST,
                                          self._stats_ref, st)
        L.check(rc)
        return idx, score, count

    def topk_device(self, queries: torch.Tensor, k: int, out=None, min_score: float = -math.inf,
                    score_mode: int = L.VM_SCORE_RAW, sum_mode: Optional[int] = None, flags: int = 0, comm=None,
                    row_offset: int = 0):
        """Device in / device out on the current stream.  -> (idx, score, count) CUDA tensors."""
        if not queries.is_cuda:
            raise TypeError("queries must be a CUDA tensor")
        if queries.device != self.device:
            raise ValueError(f"queries live on {queries.device}, the store on {self.device}")
        if queries.dim() != 2 or queries.shape[1] != self.dim:
            raise ValueError(f"expected [nq, {self.dim}] queries, got {tuple(queries.shape)}")
        if not 1 <= int(k) <= 64:
            raise ValueError(f"k = {k} outside [1, 64]")
        q = queries.contiguous()
        nq = q.shape[0]
        if out is not None:
            want = ((nq, k), torch.int64), ((nq, k), torch.float64), ((nq,), torch.int32)
            if len(out) != 3:
                raise ValueError("out must be (idx, score, count)")
            for t, (shape, dt) in zip(out, want):
                if (not isinstance(t, torch.Tensor) or t.device != self.device or t.dtype != dt
                        or tuple(t.shape) != shape or not t.is_contiguous()):
                    raise ValueError(f"out tensors must be contiguous {want} on {self.device}")
        if out is None:
            out = (torch.empty((nq, k), dtype=torch.int64, device=self.device),
                   torch.empty((nq, k), dtype=torch.float64, device=self.device),
                   torch.empty((nq,), dtype=torch.int32, device=self.device))
        idx, score, count = out
        sm = L.DEFAULT_SUM_MODE if sum_mode is None else sum_mode
        st = _stream_ptr(self.device)
        if comm is None:
            rc = self.lib.vm_topk(self._h, q.data_ptr(), _torch_dtype_code(q), L.VM_MEM_DEVICE, nq, k, float(min_score),
                                  score_mode, sm, flags, idx.data_ptr(), score.data_ptr(), count.data_ptr(),
                                  L.VM_MEM_DEVICE, self._stats_ref, st)
        else:
            rc = self.lib.vm_topk_sharded(self._h, comm.handle, int(row_offset), q.data_ptr(), _torch_dtype_code(q),
                                          L.VM_MEM_DEVICE, nq, k, float(min_score), score_mode, sm, flags,
                                          idx.data_ptr(), score.data_ptr(), count.data_ptr(), L.VM_MEM_DEVICE,
                                          self._stats_ref, st)
        L.check(rc)
        return idx, score, count


def cosine_pairs(a, b, zero_rule: int = 0, sum_mode: Optional[int] = None, device: int = 0) -> np.ndarray:
    """n independent cosines, bit-identical to PreLLMInjector._cosine_similarity (zero_rule 0,
    pre_llm_injector.py:374-388) or HybridRetriever._cosine_similarity (zero_rule 1,
    retriever_hybrid.py:655-664) on equal-length vectors; zero_rule 2 = EmbeddingUtils.cosine_similarity
    (embedding_utils.py:29-39, magnitudes through `** 0.5`; a few ulp, see include/vidmem.h)."""
    lib = L.load()
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    b = np.ascontiguousarray(np.asarray(b, dtype=np.float64))
    if a.ndim == 1:
        a, b = a[None, :], b[None, :]
    if a.shape != b.shape:
        raise ValueError("shape mismatch")
    out = np.zeros(a.shape[0], np.float64)
    sm = L.DEFAULT_SUM_MODE if sum_mode is None else sum_mode
    dev = torch.device("cuda", device)
    L.check(lib.vm_cosine_pairs(device, a.ctypes.data, b.ctypes.data, L.VM_F64, L.VM_MEM_HOST, a.shape[0], a.shape[1],
                                zero_rule, sm, out.ctypes.data, L.VM_MEM_HOST, _stream_ptr(dev)))
    return out
