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


def test_ragged_queries_and_growth_match_the_reference(monkeypatch):
    """Wrong-length, empty and Exception queries, falsy rows, and a store that outgrows its capacity several
    times: the adapter returns what the UNMODIFIED reference method returns on the same dict."""
    from hypothesis import given, settings, strategies as st
    from oracle import ref_import
    from vidmem_b200 import adapters
    import vidmem_b200.store as vstore
    monkeypatch.setattr(vstore, "EmbeddingStore", OracleBackedStore)
    d = 6
    vec = st.lists(st.integers(-3, 3).map(float), min_size=d, max_size=d)
    anyq = st.one_of(vec, st.lists(st.integers(-3, 3).map(float), min_size=0, max_size=d + 2), st.just(RuntimeError("embed failed")))
    rows = st.lists(st.one_of(vec, st.just([])), min_size=1, max_size=40)

    @settings(max_examples=40, deadline=None)
    @given(rows, st.lists(anyq, min_size=1, max_size=4), st.integers(1, 5))
    def run(row_list, queries, k):
        store = {f"u_{i}": list(r) for i, r in enumerate(row_list)}
        want = ref_import.run_batch_similarities(queries, store, k)
        res = adapters.ResidentChunkStore(initial_capacity=4)                    # forces repeated growth
        for i in range(0, len(row_list), 3):                                     # rows arrive in small insert batches
            res.upsert([(f"u_{j}", row_list[j]) for j in range(i, min(i + 3, len(row_list)))])
        assert res.topk(queries, k) == want
        assert res.ids == list(store)

    run()


def test_scalar_cosine_seams_ragged_lengths_match_the_reference(monkeypatch):
    """S2 / S4 with vectors of different lengths, zero vectors and empty vectors: the adapters' length handling
    (0.0 on mismatch for the injector; zip-truncation == zero padding for the retriever) against the reference's
    own methods.  The device kernel behind cosine_pairs is replaced by the oracle (its GPU parity is a GPU test)."""
    from hypothesis import given, settings, strategies as st
    from oracle import ref_import
    from vidmem_b200 import adapters
    import vidmem_b200.store as vstore

    def oracle_pairs(a, b, zero_rule=0, sum_mode=None, device=0):
        a, b = np.atleast_2d(np.asarray(a, np.float64)), np.atleast_2d(np.asarray(b, np.float64))
        return np.array([oracle.cosine(x, y, variant="injector" if zero_rule == 0 else "retriever") for x, y in zip(a, b)])

    monkeypatch.setattr(vstore, "cosine_pairs", oracle_pairs)
    ref = ref_import.load()
    inj = ref_import.make_injector(3)
    backend = adapters.ChunkSimilarityBackend(store=adapters.ResidentChunkStore())
    vec = st.lists(st.one_of(st.integers(-5, 5).map(float), st.floats(-1e3, 1e3, allow_nan=False, width=32)), min_size=0, max_size=9)

    @settings(max_examples=150, deadline=None)
    @given(vec, vec)
    def run(v1, v2):
        assert backend._cosine_similarity(v1, v2) == inj._cosine_similarity(v1, v2)
        assert adapters.VectorSearchBackend._cosine_similarity(v1, v2) == ref["HybridRetriever"]._cosine_similarity(v1, v2)

    run()
