"""vidmem-b200: B200-native (sm_100a) embedding-similarity engine behind the VidGraph seams.

Only the hot path named in BASELINE.json lives here: the HBM-resident embedding store, the
exact top-k scorer, the all-pairs threshold scorer and the Python adapters that keep the
reference's method signatures.  Import name: `vidmem_b200` (see vidmem_b200.py at the repo
root; the directory name carries the reference's hyphenated name).
"""
from . import _lib
from ._lib import (VM_F32, VM_BF16, VM_F64, VM_SCORE_RAW, VM_SCORE_NEO4J, VM_SUM_NAIVE, VM_SUM_NEUMAIER,
                   VM_FLAG_ASYNC, VM_FLAG_FORCE_EXACT, VM_FLAG_FORCE_SIMT, VM_FLAG_FORCE_TC, VM_FLAG_TIMING, VM_FLAG_NO_SPLIT, VM_FLAG_SPLIT, VidmemError)

__all__ = ["_lib", "EmbeddingStore", "cosine_pairs", "VidmemError"]


def __getattr__(name):
    # torch-dependent pieces are imported lazily so that `import vidmem_b200` stays cheap
    if name in ("EmbeddingStore", "cosine_pairs"):
        from . import store
        return getattr(store, name)
    if name in ("adapters", "sharded", "dedup", "store"):
        import importlib
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)
