"""Call-surface compatibility on the UNMODIFIED reference classes (CPU, authoring container only: skipped when
/root/reference is absent).  The device store is replaced by a test double that answers through the oracle, so
what is exercised here is the seam itself: method rebinding, coroutine signature, id <-> row bookkeeping, return
types and the error convention -- compared with what the reference's own method returns on the same inputs."""
import asyncio
import os

import numpy as np
import pytest

from oracle import oracle, synth

pytestmark = pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference checkout not present")


from doubles import OracleBackedStore


def test_install_on_real_reference_injector(monkeypatch):
    from oracle import ref_import
    from vidmem_b200 import adapters
    import vidmem_b200.store as vstore
    monkeypatch.setattr(vstore, "EmbeddingStore", OracleBackedStore)
    n, d, k = 120, 64, 3
    X = synth.synth_rows(7, 0, n, d)
    store = {f"u_{i // 4}_{i % 4}": [float(v) for v in X[i]] for i in range(n)}
    store["u_bad"] = []                                                   # falsy embedding is skipped (:363)
    Q = synth.synth_queries(8, 4, d, 7, n)
    queries = [[float(v) for v in q] for q in Q] + [RuntimeError("embed failed")]

    # what the reference itself returns
    want = ref_import.run_batch_similarities(queries, store, k)

    # the same, through the adapter bound onto an unmodified PreLLMInjector instance
    inj = ref_import.make_injector(k)

    async def fake_get(_handler):
        return store

    inj._get_chunk_embeddings = fake_get                                   # the reference's own fetch seam (:390-412)
    backend = adapters.install_injector(inj)
    assert asyncio.iscoroutinefunction(inj._calculate_batch_similarities)
    got = asyncio.run(inj._calculate_batch_similarities(queries, object()))
    assert got == want and got[-1] == []
    assert all(isinstance(cid, str) and isinstance(s, float) for lst in got for cid, s in lst)
    assert len(backend.store) == n + 1 and backend.store.ids[-1] == "u_bad"
