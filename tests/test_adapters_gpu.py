"""GPU tests of the reference-facing adapters (S1-S6) against the oracle; the reference objects are
duck-typed stand-ins because /root/reference does not exist on the GPU box."""
import asyncio
import types

import numpy as np
import pytest

from oracle import oracle, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ad():
    import torch
    assert torch.cuda.is_available()
    from vidmem_b200 import adapters
    return adapters


class FakeInjector:
    """Duck-typed PreLLMInjector: only what S1 touches (embedder_config, _get_chunk_embeddings)."""

    def __init__(self, store_dict, k, k2=2):
        self.embedder_config = types.SimpleNamespace(top_k_chunk_with_batch_similarity=k, top_k_similar_batch=k2)
        self.store_dict = store_dict
        self.fetches = 0

    async def _get_chunk_embeddings(self, neo4j_handler):
        self.fetches += 1
        return self.store_dict


def _lists(a):
    return [[float(v) for v in r] for r in a]


def test_s1_mirror_mode_matches_reference_semantics(ad):
    n, d, k = 300, 96, 3
    X = synth.synth_rows(71, 0, n, d)
    X[10] = X[4]
    row_ok = np.ones(n, np.uint8); row_ok[[7, 200]] = 0
    store = {f"c{i}": (_lists(X[i:i + 1])[0] if row_ok[i] else (None if i == 7 else [])) for i in range(n)}
    Q = synth.synth_queries(72, 5, d, 71, n)
    Q[1] = X[4]
    queries = _lists(Q)
    queries[3] = RuntimeError("embedding failed")
    queries.append([0.5] * (d - 1))                                  # wrong length -> every score 0.0
    inj = FakeInjector(store, k)
    backend = ad.install_injector(inj, initial_capacity=64)          # forces the store to grow
    got = asyncio.run(inj._calculate_batch_similarities(queries, object()))
    qok = np.array([1, 1, 1, 0, 1], np.uint8)
    ref = oracle.batch_similarities(Q, X, k, query_ok=qok, row_ok=row_ok)
    assert len(got) == 6
    for gi, ri in zip(got[:5], ref):
        assert gi == [(f"c{r}", s) for r, s in ri]
    assert got[3] == []
    assert got[5] == [("c0", 0.0), ("c1", 0.0), ("c2", 0.0)]
    # second call: the dict grew -> only the new rows are appended, results follow
    X2 = synth.synth_rows(73, 0, 20, d)
    for i in range(20):
        store[f"n{i}"] = _lists(X2[i:i + 1])[0]
    got2 = asyncio.run(inj._calculate_batch_similarities(_lists(X2[3:4]), object()))
    assert got2[0][0][0] == "n3" and got2[0][0][1] == oracle.cosine(X2[3], X2[3])
    assert inj.fetches == 2 and len(backend.store) == n + 20
    assert inj._cosine_similarity(queries[0], store["c5"]) == oracle.cosine(Q[0], X[5])
    assert inj._cosine_similarity([1.0, 2.0], [1.0]) == 0.0


def test_s6_insert_hook_and_merge(ad):
    d, k, k2 = 64, 3, 2
    X = synth.synth_rows(81, 0, 120, d)
    backend = ad.ChunkSimilarityBackend(mirror_fetch=False, initial_capacity=256)
    chunks = [{"id": f"u_{i // 4}_{i % 4}", "content": f"text {i}", "index": i, "embedding": _lists(X[i:i + 1])[0]}
              for i in range(120)]
    backend.on_chunks_inserted(chunks[:50])
    backend.on_chunks_inserted(chunks[50:])
    backend.on_chunks_inserted([{"id": "u_0_1", "content": "new", "embedding": _lists(X[99:100])[0]}])  # upsert by id
    X2 = X.copy(); X2[1] = X[99]
    Q = synth.synth_queries(82, 4, d, 81, 120); Q[1] = Q[0]
    inj = FakeInjector({}, k, k2)
    got = asyncio.run(backend._calculate_batch_similarities(inj, _lists(Q), object()))
    ref = oracle.batch_similarities(Q, X2, k)
    ids = [c["id"] for c in chunks]
    assert got == [[(ids[r], s) for r, s in lst] for lst in ref]
    assert inj.fetches == 0                                          # resident mode never fetches
    merged = backend.merge_top_similar(got, k2)
    assert merged == [(ids[r], s) for r, s in oracle.merge_max_by_id(ref, k2)]
    # f2: both steps in one device pass, and the fallback for queries that need the adapter's special cases
    lists, seeds = backend.similarities_and_top_similar(_lists(Q), k, k2)
    assert lists == got and seeds == merged
    odd = _lists(Q) + [RuntimeError("embed failed"), []]
    lists2, seeds2 = backend.similarities_and_top_similar(odd, k, k2)
    assert lists2[:4] == got and lists2[4] == [] and seeds2 == backend.merge_top_similar(lists2, k2)


def test_s3_vector_search_and_s4_filter(ad):
    d = 48
    X = synth.synth_rows(91, 0, 80, d)
    store = ad.ResidentChunkStore(initial_capacity=128)
    store.upsert(((f"c{i}", _lists(X[i:i + 1])[0]) for i in range(80)),
                 meta={f"c{i}": {"content": f"chunk {i}", "time": float(i)} for i in range(80)})
    q = X[9].copy(); q[::5] = 0.25

    class Emb:
        async def aembed_query(self, text):
            return [float(v) for v in q]

    retr = types.SimpleNamespace(config=types.SimpleNamespace(top_k_chunks=4),
                                 neo4j_handler=types.SimpleNamespace(embedder=Emb()))
    backend = ad.install_retriever(retr, store)
    got = asyncio.run(retr._vector_search_chunks(None, "what happened?"))
    ref = oracle.vector_search(q, X, 4, 0.3)
    assert [(g["id"], g["score"]) for g in got] == [(f"c{r}", s) for r, s in ref]
    assert got[0]["source"] == "vector" and got[0]["content"] == "chunk 9" and got[0]["time"] == 9.0
    thr = oracle.cosine(q, X[3], "retriever")
    kept = backend.filter_segments(list(q), _lists(X[:12]), thr, 50)
    assert kept == oracle.threshold_filter_ge(q, X[:12], thr, 50)
    assert (3, thr) in kept
    a, b = [0.3, -0.2, 0.9, 0.4], [0.1, 0.7]
    assert backend._cosine_similarity(a, b) == oracle.cosine(a, b, "retriever")   # zip-truncated dot


def test_f4_rerank_prefilter_is_opt_in(ad):
    """install_retriever(..., rerank_prefilter=N): the reference's own _rerank_chunks receives only the N chunks closest
    to the query by the resident embeddings (original order kept, unknown ids always kept); without the option, or with
    the reranker disabled, nothing changes."""
    d = 32
    X = synth.synth_rows(71, 0, 40, d)
    store = ad.ResidentChunkStore(initial_capacity=64)
    store.upsert((f"c{i}", _lists(X[i:i + 1])[0]) for i in range(40))
    q = X[5].copy(); q[::3] = -0.5

    class Emb:
        async def aembed_query(self, text):
            return [float(v) for v in q]

    def make(use):
        seen = []

        class Retr:
            config = types.SimpleNamespace(top_k_chunks=4, use_reranker=use)
            neo4j_handler = types.SimpleNamespace(embedder=Emb())

            async def _rerank_chunks(self, query, chunks, raise_on_failure=False):
                seen.append([c["id"] for c in chunks])
                return list(reversed(chunks))

            async def _vector_search_chunks(self, session, query):
                return []
        return Retr(), seen

    chunks = [{"id": f"c{i}", "content": str(i)} for i in (30, 5, 17, 2, 9, 21)] + [{"id": "not-resident", "content": "x"}]
    scores = {i: oracle.cosine(q, X[i], "retriever") for i in (30, 5, 17, 2, 9, 21)}
    want = sorted(scores, key=lambda i: -scores[i])[:2]
    retr, seen = make(True)
    ad.install_retriever(retr, store, rerank_prefilter=3)
    out = asyncio.run(retr._rerank_chunks("q", chunks))
    assert seen == [[c["id"] for c in chunks if c["id"] == "not-resident" or int(c["id"][1:]) in want]] and len(seen[0]) == 3
    assert [c["id"] for c in out] == list(reversed(seen[0]))
    retr, seen = make(True)
    ad.install_retriever(retr, store)                                # not asked for: the reference method is untouched
    asyncio.run(retr._rerank_chunks("q", chunks))
    assert seen == [[c["id"] for c in chunks]]
    retr, seen = make(False)
    ad.install_retriever(retr, store, rerank_prefilter=3)            # reranker disabled: chunks pass through unfiltered
    asyncio.run(retr._rerank_chunks("q", chunks))
    assert seen == [[c["id"] for c in chunks]]


def test_s5_representative(ad):
    E = synth.synth_rows(3, 0, 6, 768, dup_period=2)
    g = types.SimpleNamespace(embedding_model=types.SimpleNamespace(encode=lambda s: np.asarray(s, np.float32)))
    backend = ad.PruneBackend()
    assert backend._get_representative_relation(g, E) == oracle.representative(E)[0]
    assert backend._are_same_context(g, E[:1], 0.8) is False


def test_ids_before_first_embedding_and_boundaries(ad):
    """Chunks whose embeddings failed arrive first: the dimension is not known yet, yet store order holds."""
    store = ad.ResidentChunkStore(initial_capacity=8)
    store.upsert([("a", None), ("b", [])])
    assert len(store) == 2 and store.store is None
    assert store.topk([[1.0, 2.0]], 3) == [[]]                       # nothing scorable yet
    X = synth.synth_rows(3, 0, 20, 16)
    store.upsert((f"c{i}", [float(v) for v in X[i]]) for i in range(20))
    assert store.ids[:3] == ["a", "b", "c0"] and len(store) == 22
    got = store.topk([[float(v) for v in X[4]]], 3)[0]
    ref = oracle.batch_similarities(X[4:5], X, 3)[0]
    assert got == [(f"c{r}", s) for r, s in ref]
    store.upsert([("a", [float(v) for v in X[4]])])                 # the failed chunk is re-embedded later (MERGE by id)
    got = store.topk([[float(v) for v in X[4]]], 2)[0]
    assert [g[0] for g in got] == ["a", "c4"] and got[0][1] == got[1][1] == oracle.cosine(X[4], X[4])


def test_f1_hydrate_and_f3_sidecar_roundtrip(ad, tmp_path):
    """Bulk hydration from a (fake) Neo4j session, then save -> load of the resident store: same ids, bit-identical
    answers; and re-hydration from a JSON graph export."""
    import json
    from test_hydrate_cpu import _Handler, _records
    n, d, k = 700, 64, 5
    X, recs = _records(n, d)
    backend = ad.ChunkSimilarityBackend(initial_capacity=64)               # forces several growth steps
    assert asyncio.run(backend.hydrate(_Handler(recs), page_rows=128)) == n - 2
    kept = [r for r in recs if isinstance(r["embedding"], list) and r["chunk_id"]]
    Q = _lists(synth.synth_queries(6, 4, d, 5, n))
    want = oracle.batch_similarities(np.array(Q), np.array([r["embedding"] if r["embedding"] else [0.0] * d for r in kept]), k,
                                     row_ok=np.array([1 if r["embedding"] else 0 for r in kept], np.uint8))
    want = [[(kept[r]["chunk_id"], s) for r, s in lst] for lst in want]
    assert backend.store.topk(Q, k) == want
    assert backend.store.dtype == "f64" and backend.store.store.exact and backend.store.store.growable   # the adapters' default
    for dt in ("f64", "bf16", "f32", "f64+bf16"):
        src = backend.store if dt == "f64" else ad.ResidentChunkStore(dt)
        if dt != "f64":
            src.upsert([(r["chunk_id"], r["embedding"]) for r in kept])
        p = str(tmp_path / f"chunks_{dt}")
        src.save(p)
        back = ad.ResidentChunkStore.load(p)
        assert back.ids == src.ids and back.dtype == dt and back.meta == src.meta
        assert back.topk(Q, k) == src.topk(Q, k)
        fresh = _lists(synth.synth_rows(99, 0, 1, d))
        back.upsert([("late", fresh[0])])                                  # still growable after a load
        assert back.topk(fresh, 1)[0][0][0] == "late"
    nodes = [{"name": None, "labels": ["Chunk"], "properties": {"id": r["chunk_id"], "content": r["content"], "embedding": r["embedding"]}}
             for r in kept]
    path = tmp_path / "export.json"
    path.write_text(json.dumps({"graph_uuid": "g", "nodes": nodes, "relationships": []}))
    st = ad.ResidentChunkStore()
    assert st.load_export(str(path)) == len(kept) and st.topk(Q, k) == want
