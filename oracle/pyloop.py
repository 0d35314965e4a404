"""Interpreter-level restatement of the reference's scoring loop -- TEST INFRASTRUCTURE ONLY (like everything under
oracle/): used by tests/ to pin the C restatement (vm_oracle.c) to what CPython itself computes, and by bench.py's
cpu_baseline leg to time the reference's OWN execution model (one interpreter thread, generator expressions over Python
floats) on a bounded sample -- the C port in vm_oracle.c is 2-3 orders of magnitude faster than the code it restates.

Follows src/components/pre_llm_injector.py:346-388: per query, score every store row whose embedding is truthy, stable
sort by score descending, keep the first top_k; Exception queries give []; a length mismatch or a zero norm scores 0.0.
The summation order is whatever the running interpreter's sum() does (Neumaier-compensated from CPython 3.12 on)."""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple


def cosine(u: Sequence[float], v: Sequence[float]) -> float:
    """pre_llm_injector.py:374-388."""
    if len(u) != len(v):
        return 0.0
    dot = sum(x * y for x, y in zip(u, v))
    nu = math.sqrt(sum(x * x for x in u))
    nv = math.sqrt(sum(y * y for y in v))
    if nu == 0 or nv == 0:
        return 0.0
    return dot / (nu * nv)


def batch_similarities(queries: Sequence, store: Dict[str, List[float]], top_k: int) -> List[List[Tuple[str, float]]]:
    """pre_llm_injector.py:346-372 with the store dict passed in (the reference fetches it from Neo4j at :353)."""
    out: List[List[Tuple[str, float]]] = []
    for q in queries:
        if isinstance(q, Exception):
            out.append([])
            continue
        scored = [(cid, float(cosine(q, emb))) for cid, emb in store.items() if emb]
        scored.sort(key=lambda t: t[1], reverse=True)          # stable: ties keep dict order
        out.append(scored[:top_k])
    return out
