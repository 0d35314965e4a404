"""Rows f1 / f3 on CPU: bulk hydration from a (fake) Neo4j session and re-hydration from a graph export.
The device store is replaced by the oracle-backed double; what is tested is the host logic: the Cypher
text, the reference's row filter, paging, row order, and the id table."""
import asyncio
import json

import numpy as np

from doubles import OracleBackedStore
from oracle import oracle, synth


class _Result:
    def __init__(self, records):
        self.records = records

    def __aiter__(self):
        async def gen():
            for r in self.records:
                yield r
        return gen()


class _Session:
    def __init__(self, handler):
        self.h = handler

    async def __aenter__(self):
        return self

    async def __aexit__(self, *a):
        return False

    async def run(self, query, **params):
        self.h.queries.append((query, params))
        recs = self.h.records
        if "LIMIT" in query:
            recs = recs[:int(query.split("LIMIT")[1].split()[0])]
        return _Result(recs)


class _Handler:
    run_uuid = "uuid-1"

    def __init__(self, records):
        self.records, self.queries = records, []
        self.driver = self

    def session(self):
        return _Session(self)


def _records(n, d):
    X = synth.synth_rows(5, 0, n, d)
    recs = [{"chunk_id": f"u_{i}", "embedding": [float(v) for v in X[i]], "content": f"text {i}"} for i in range(n)]
    recs[3]["embedding"] = "not-a-list"        # dropped by the reference's isinstance(list) filter (:405)
    recs[4]["chunk_id"] = ""                   # dropped: falsy id
    recs[6]["embedding"] = []                  # kept as a row, skipped when scoring (:363)
    return X, recs


def test_hydrate_pages_and_matches_reference_filter(monkeypatch):
    import vidmem_b200.store as vstore
    from vidmem_b200 import adapters
    monkeypatch.setattr(vstore, "EmbeddingStore", OracleBackedStore)
    n, d = 57, 32
    X, recs = _records(n, d)
    h = _Handler(recs)
    backend = adapters.ChunkSimilarityBackend()
    got = asyncio.run(backend.hydrate(h, page_rows=10))
    kept = [r for r in recs if isinstance(r["embedding"], list) and r["chunk_id"]]
    assert got == len(kept) == n - 2 and backend.mirror_fetch is False
    assert backend.store.ids == [r["chunk_id"] for r in kept]
    q, params = h.queries[0]
    assert "MATCH (c:Chunk:GraphNode)" in q and "c.embedding IS NOT NULL" in q and "LIMIT" not in q
    assert params == {"graph_uuid": "uuid-1"}
    assert backend.store.meta["u_7"]["content"] == "text 7"
    # answers == the reference formula over the same rows
    Q = synth.synth_queries(6, 3, d, 5, n)
    want = oracle.batch_similarities(Q, np.array([r["embedding"] if r["embedding"] else [0.0] * d for r in kept]), 3,
                                     row_ok=np.array([1 if r["embedding"] else 0 for r in kept], np.uint8))
    res = backend.store.topk([[float(v) for v in q] for q in Q], 3)
    assert res == [[(kept[r]["chunk_id"], s) for r, s in lst] for lst in want]
    # the reference's cap stays available behind the flag
    h2 = _Handler(recs)
    b2 = adapters.ChunkSimilarityBackend()
    assert asyncio.run(b2.hydrate(h2, limit=20)) == 18 and "LIMIT 20" in h2.queries[0][0]


def test_load_export_takes_chunk_nodes_in_file_order(monkeypatch, tmp_path):
    import vidmem_b200.store as vstore
    from vidmem_b200 import adapters
    monkeypatch.setattr(vstore, "EmbeddingStore", OracleBackedStore)
    X = synth.synth_rows(9, 0, 6, 16)
    nodes = [{"name": None, "labels": ["Chunk"], "properties": {"id": f"c{i}", "content": f"t{i}", "embedding": [float(v) for v in X[i]]}}
             for i in range(6)]
    nodes.insert(2, {"name": "alice", "labels": ["Person"], "properties": {"id": "p1", "embedding": [1.0] * 16}})   # not a Chunk
    nodes.insert(4, {"name": None, "labels": ["Chunk"], "properties": {"id": "c_noemb", "content": "x"}})           # no embedding
    path = tmp_path / "export.json"
    path.write_text(json.dumps({"graph_uuid": "g", "nodes": nodes, "relationships": [], "export_format_version": "1.0"}))
    st = adapters.ResidentChunkStore()
    assert st.load_export(str(path)) == 6
    assert st.ids == [f"c{i}" for i in range(6)] and st.meta["c3"]["content"] == "t3"
    assert np.array_equal(st.store.X, X)                                    # default store dtype: binary64, values as given


def test_s3_vector_search_adapter_shape_threshold_and_error_convention(monkeypatch):
    """S3 on CPU (device store replaced by the oracle): result dict keys, Neo4j-normalised score with the strict
    > 0.3 filter, best first, content/time from the insert hook's metadata, [] on any error (:321-323)."""
    import types
    import vidmem_b200.store as vstore
    from vidmem_b200 import adapters
    monkeypatch.setattr(vstore, "EmbeddingStore", OracleBackedStore)
    n, d = 40, 16
    X = synth.synth_rows(3, 0, n, d)
    backend = adapters.ChunkSimilarityBackend(mirror_fetch=False)
    backend.on_chunks_inserted([{"id": f"u_{i}", "content": f"text {i}", "time": f"00:{i:02d}", "embedding": [float(v) for v in X[i]]}
                                for i in range(n)])
    q = [float(v) for v in X[7]]

    class Embedder:
        async def aembed_query(self, text):
            if text == "boom":
                raise RuntimeError("embedding service down")
            return q

    retriever = types.SimpleNamespace(neo4j_handler=types.SimpleNamespace(embedder=Embedder()), config=types.SimpleNamespace(top_k_chunks=5))
    vs = adapters.install_retriever(retriever, backend.store)
    got = asyncio.run(retriever._vector_search_chunks(None, "what happened?"))
    want = oracle.vector_search(np.array(q), X, 5, min_score=0.3)
    assert [g["id"] for g in got] == [f"u_{r}" for r, _ in want] and [g["score"] for g in got] == [s for _, s in want]
    assert got[0]["id"] == "u_7" and got[0]["score"] == 1.0 and got[0]["content"] == "text 7" and got[0]["time"] == "00:07"
    assert all(set(g) == {"id", "time", "content", "score", "source"} and g["source"] == "vector" and g["score"] > 0.3 for g in got)
    assert asyncio.run(retriever._vector_search_chunks(None, "boom")) == []
    assert isinstance(vs, adapters.VectorSearchBackend)
