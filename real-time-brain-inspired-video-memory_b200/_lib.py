"""ctypes binding of libvidmem.so (include/vidmem.h).

There is no CPU fallback: if the shared library is missing this module raises at import
(`VIDMEM_BUILD=1` builds it first with nvcc), and every compute entry point fails loudly
without an sm_100 device.
"""
from __future__ import annotations

import ctypes as C
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VIDMEM_LIB", os.path.join(HERE, "libvidmem.so"))  # override: A/B builds

# enums (include/vidmem.h)
VM_OK = 0
VM_ERR_BADARG, VM_ERR_OOM, VM_ERR_CUDA, VM_ERR_NCCL, VM_ERR_OVERFLOW, VM_ERR_UNSUPPORTED, VM_ERR_STATE = -1, -2, -3, -4, -5, -6, -7
VM_F32, VM_BF16, VM_F64 = 0, 1, 2
VM_MEM_HOST, VM_MEM_DEVICE = 0, 1
VM_SCORE_RAW, VM_SCORE_NEO4J = 0, 1
VM_SUM_NAIVE, VM_SUM_NEUMAIER = 0, 1
VM_FLAG_ASYNC, VM_FLAG_FORCE_EXACT, VM_FLAG_FORCE_SIMT, VM_FLAG_FORCE_TC, VM_FLAG_TIMING, VM_FLAG_NO_SPLIT, VM_FLAG_SPLIT = 1, 2, 4, 8, 16, 32, 64

#: summation order CPython's builtin sum() uses in THIS interpreter (what the reference would compute here)
DEFAULT_SUM_MODE = VM_SUM_NEUMAIER if sys.version_info >= (3, 12) else VM_SUM_NAIVE

EXPORTS = [
    "vm_version", "vm_last_error", "vm_device_info", "vm_ld",
    "vm_store_create", "vm_store_attach", "vm_store_create_exact", "vm_store_attach_exact", "vm_store_create_growable", "vm_store_reserve",
    "vm_store_max_capacity", "vm_store_resident_bytes", "vm_store_rows_ptr", "vm_store_inv_norms_ptr", "vm_store_rows_exact_ptr", "vm_store_destroy", "vm_store_size", "vm_store_capacity", "vm_store_dim",
    "vm_store_ld", "vm_store_append", "vm_store_update", "vm_store_invalidate", "vm_store_set_size", "vm_store_clear",
    "vm_store_last_scan_ms",
    "vm_store_avg_scan_ms", "vm_store_read_counters", "vm_store_band_keys",
    "vm_topk", "vm_topk_sharded", "vm_merge_topk_lists", "vm_topk_packed_bytes", "vm_merge_topk_packed", "vm_merge_max_by_id", "vm_cosine_pairs", "vm_pairs_above", "vm_pairs_above_sharded",
    "vm_comm_unique_id", "vm_comm_init_rank", "vm_comm_destroy", "vm_comm_exchange_bytes", "vm_comm_attach_peer_buffers", "vm_comm_nranks", "vm_comm_rank", "vm_synth_fill",
]


class TopkStats(C.Structure):
    _fields_ = [("scan_kernel", C.c_int32), ("scan_launches", C.c_int32), ("uncertified", C.c_int32),
                ("candidates", C.c_int32), ("scan_ctas", C.c_int32), ("scan_stages", C.c_int32), ("full_rescans", C.c_int32), ("scan_variant", C.c_int32)]


class StoreCounters(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("batches", "queries", "uncertified", "band_settled", "collect_settled", "full_rescans", "bound_violations")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class VidmemError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libvidmem error {code}: {msg}")
        self.code = code


_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if os.environ.get("VIDMEM_BUILD") == "1":
            from . import build as _build
            _build.build()
        else:
            raise ImportError(f"{LIB_PATH} is missing: build it with `python {HERE}/build.py` "
                              "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, i64, i32, dbl = C.c_void_p, C.c_int64, C.c_int32, C.c_double
    ci, u64, sz = C.c_int, C.c_uint64, C.c_size_t
    P = C.POINTER
    sig = {
        "vm_version": (ci, []),
        "vm_last_error": (C.c_char_p, []),
        "vm_device_info": (ci, [ci, P(ci), P(ci), P(ci), P(sz)]),
        "vm_ld": (ci, [ci]),
        "vm_store_create": (ci, [P(vp), ci, ci, ci, i64]),
        "vm_store_attach": (ci, [P(vp), ci, ci, ci, i64, vp, vp]),
        "vm_store_create_exact": (ci, [P(vp), ci, ci, ci, i64]),
        "vm_store_attach_exact": (ci, [P(vp), ci, ci, ci, i64, vp, vp, vp]),
        "vm_store_create_growable": (ci, [P(vp), ci, ci, ci, ci, i64, i64]),
        "vm_store_reserve": (ci, [vp, i64]),
        "vm_store_max_capacity": (i64, [vp]),
        "vm_store_resident_bytes": (sz, [vp]),
        "vm_store_rows_ptr": (vp, [vp]),
        "vm_store_inv_norms_ptr": (vp, [vp]),
        "vm_store_rows_exact_ptr": (vp, [vp]),
        "vm_store_destroy": (ci, [vp]),
        "vm_store_size": (i64, [vp]),
        "vm_store_capacity": (i64, [vp]),
        "vm_store_dim": (ci, [vp]),
        "vm_store_ld": (ci, [vp]),
        "vm_store_append": (ci, [vp, vp, ci, ci, i64, P(i64), vp]),
        "vm_store_update": (ci, [vp, i64, vp, ci, ci, i64, vp]),
        "vm_store_invalidate": (ci, [vp, vp, i64, vp]),
        "vm_store_set_size": (ci, [vp, i64, i64, vp]),
        "vm_store_clear": (ci, [vp]),
        "vm_store_last_scan_ms": (ci, [vp, P(C.c_float)]),
        "vm_store_avg_scan_ms": (ci, [vp, P(C.c_float), P(C.c_int)]),
        "vm_store_read_counters": (ci, [vp, P(StoreCounters), ci]),
        "vm_store_band_keys": (ci, [vp, ci, P(i64), P(i64), P(ci)]),
        "vm_topk": (ci, [vp, vp, ci, ci, ci, ci, dbl, ci, ci, ci, vp, vp, vp, ci, P(TopkStats), vp]),
        "vm_topk_sharded": (ci, [vp, vp, i64, vp, ci, ci, ci, ci, dbl, ci, ci, ci, vp, vp, vp, ci, P(TopkStats), vp]),
        "vm_merge_topk_lists": (ci, [ci, vp, vp, vp, ci, ci, ci, vp, vp, vp, vp]),
        "vm_topk_packed_bytes": (sz, [ci, ci]),
        "vm_merge_topk_packed": (ci, [ci, vp, ci, ci, ci, vp, vp, vp, vp]),
        "vm_merge_max_by_id": (ci, [ci, vp, vp, vp, ci, ci, ci, vp, vp, vp, vp]),
        "vm_cosine_pairs": (ci, [ci, vp, vp, ci, ci, i64, ci, ci, ci, vp, ci, vp]),
        "vm_pairs_above": (ci, [ci, vp, ci, i64, ci, C.c_float, i64, vp, vp, vp, vp, ci, ci, ci, vp]),
        "vm_pairs_above_sharded": (ci, [vp, vp, ci, i64, ci, C.c_float, i64, vp, vp, vp, vp, ci, ci, vp]),
        "vm_comm_unique_id": (ci, [vp]),
        "vm_comm_init_rank": (ci, [P(vp), ci, ci, ci, vp]),
        "vm_comm_destroy": (ci, [vp]),
        "vm_comm_exchange_bytes": (sz, []),
        "vm_comm_attach_peer_buffers": (ci, [vp, P(vp), ci]),
        "vm_comm_nranks": (ci, [vp]),
        "vm_comm_rank": (ci, [vp]),
        "vm_synth_fill": (ci, [ci, vp, ci, u64, i64, i64, ci, u64, vp]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)  # AttributeError if the .so does not export a declared symbol
        f.restype, f.argtypes = res, args
    if L.vm_version() != 1:
        raise ImportError("libvidmem ABI version mismatch")
    _lib = L
    return L


def check(rc: int, allow=()) -> int:
    if rc != VM_OK and rc not in allow:
        msg = load().vm_last_error()
        raise VidmemError(rc, msg.decode("utf-8", "replace") if msg else "")
    return rc
