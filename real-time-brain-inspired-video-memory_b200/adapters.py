"""Drop-in adapters: the reference's Python seams backed by libvidmem.

Each adapter keeps the signature, return shape and error convention of the reference method it
replaces (file:line cited per method) and can be bound onto an unmodified reference object
with `install_*`.  Row identity: chunk ids are strings in the reference
("{run_uuid}_{batch_idx}_{i}", src/components/pre_llm_injector.py:91); the engine works on dense
row indices, so the adapters keep the id <-> row table on the host.  Row index == first-seen
order == the reference's dict insertion order, which is what its stable sort uses to break ties
(SURVEY.md 9.2).
"""
from __future__ import annotations

import math
import types
from typing import Any, Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib as L


def _falsy_embedding(emb) -> bool:
    # `if existing_emb:` (src/components/pre_llm_injector.py:363): None and [] are skipped
    return emb is None or (hasattr(emb, "__len__") and len(emb) == 0)


class ResidentChunkStore:
    """Host-side id table + one HBM-resident EmbeddingStore that grows IN PLACE: the store reserves virtual
    addresses for `max_capacity` rows (default: what the GPU's memory could hold) and backs them with HBM as chunks
    arrive -- resident rows are never copied, there is no second allocation.

    Replaces the dict returned by PreLLMInjector._get_chunk_embeddings
    (src/components/pre_llm_injector.py:390-412) and is fed by the insert hook S6
    (src/components/neo4j_handler.py:221-253)."""

    def __init__(self, dtype: str = "f64", device: int = 0, initial_capacity: int = 8192,
                 max_capacity: Optional[int] = None):
        # dtype "f64" (default): a binary64 store -- the reference scores Python floats, so the drop-in keeps the rows
        # exactly as given and its results equal the reference's for ANY input; "f64+bf16" halves the scan traffic,
        # "f32" / "bf16" store rounded values (exact for the stored values; 3x / 6x less HBM per row)
        self.dtype, self.device = dtype, device
        self.initial_capacity = int(initial_capacity)
        self.max_capacity = max_capacity
        self.store = None                      # created lazily: the dimension is whatever the embedder returns
        self.dim: Optional[int] = None
        self.ids: List[str] = []               # row -> chunk id
        self.row_of: Dict[str, int] = {}       # chunk id -> row
        self.meta: Dict[str, Dict[str, Any]] = {}  # chunk id -> {"content", "time"} for vector search results
        self._mirror_prev: Optional[Dict[str, Any]] = None   # last dict handed to sync_from_dict (mirror mode)
        #: True while the resident rows are known to be ALL chunks of the graph (fed by the insert hook from an
        #: empty graph, or hydrated without a LIMIT).  A mirror of _get_chunk_embeddings is not: that query is
        #: capped at 5000 rows and only re-read at ingest (pre_llm_injector.py:394-399), whereas the Cypher scan of
        #: _vector_search_chunks sees every chunk -- VectorSearchBackend serves searches from HBM only when True.
        self.complete = True

    def __len__(self) -> int:
        return len(self.ids)

    # -- storage ---------------------------------------------------------------------------------
    def _ensure(self, dim: int, extra: int) -> None:
        from .store import EmbeddingStore
        if self.store is None:
            self.dim = int(dim)
            cap = max(self.initial_capacity, 2 * extra)
            top = self.max_capacity if self.max_capacity is not None else self._device_row_limit(self.dim)
            self.store = EmbeddingStore(self.dim, cap, self.dtype, self.device, max_capacity=max(int(top), cap))
            return
        if dim != self.dim:
            raise ValueError(f"embedding dimension changed from {self.dim} to {dim}")
        need = len(self.ids) + extra
        if need > self.store.capacity:
            if need > self.store.max_capacity:
                raise MemoryError(f"{need} rows exceed the store's reserved maximum of {self.store.max_capacity}")
            # back more of the reserved range: no reallocation, no copy, row addresses unchanged
            self.store.reserve(min(max(2 * self.store.capacity, need), self.store.max_capacity))

    def _device_row_limit(self, dim: int) -> int:
        """Rows the device's memory could hold at most (virtual addresses are reserved for that many)."""
        from .store import device_row_limit
        return device_row_limit(dim, self.dtype, self.device)

    def clear(self) -> None:
        if self.store is not None:
            self.store.clear()
        self.ids, self.row_of = [], {}
        self._mirror_prev = None
        self.complete = True

    def upsert(self, items: Iterable[Tuple[str, Any]], meta: Optional[Dict[str, Dict[str, Any]]] = None) -> None:
        """Insert hook: (chunk_id, embedding) pairs.  Unknown ids are appended, known ids are
        overwritten in place (the reference MERGEs by id, neo4j_handler.py:229).  A falsy embedding
        keeps its row slot but is marked skipped."""
        # an id written twice in one batch is ONE row: first-seen position, last value (dict semantics, like the
        # reference's {id: embedding} and its MERGE-by-id inserts)
        merged: Dict[str, Any] = {}
        for cid, emb in items:
            merged[cid] = emb
        items = list(merged.items())
        new_rows, new_ids, invalid = [], [], []
        if self.dim is not None:
            dim = self.dim                                    # fixed once the store exists: other lengths score 0.0
        else:
            # first embeddable batch: the store dimension is the most common length of the batch (ties -> the
            # earliest), so one malformed vector cannot fix a wrong dimension for good
            lens = [len(e) for _, e in items if not _falsy_embedding(e)]
            dim = max(sorted(set(lens), key=lens.index), key=lens.count) if lens else None
        if dim is None:
            # no embeddable vector seen yet, so the dimension is still unknown: only remember the ids (they
            # keep their place in store order and become skipped rows once the store exists)
            for cid, _ in items:
                if cid not in self.row_of:
                    self.row_of[cid] = len(self.ids)
                    self.ids.append(cid)
            if meta:
                self.meta.update(meta)
            return
        if self.store is None and self.ids:
            # ids registered before the dimension was known: materialise them as skipped rows
            self._ensure(dim, len(self.ids) + len(items))
            self.store.append(np.zeros((len(self.ids), dim), np.float64))
            self.store.invalidate(range(len(self.ids)))
        for cid, emb in items:
            if cid in self.row_of:
                row = self.row_of[cid]
                if _falsy_embedding(emb):
                    invalid.append(row)
                elif len(emb) == dim:
                    self._ensure(dim, 0)
                    self.store.update(row, np.asarray(emb, dtype=np.float64)[None, :])
                else:
                    self._ensure(dim, 0)
                    self.store.update(row, np.zeros((1, dim), np.float64))        # length mismatch scores 0.0 (:378-379)
            else:
                new_ids.append(cid)
                new_rows.append(None if _falsy_embedding(emb) else np.asarray(emb, dtype=np.float64))
        if new_ids:
            self._ensure(dim, len(new_ids))
            block = np.zeros((len(new_ids), self.dim), np.float64)
            for i, r in enumerate(new_rows):
                if r is None:
                    invalid.append(len(self.ids) + i)
                elif len(r) != self.dim:
                    # the reference scores a length mismatch as 0.0 (:378-379) == a zero row
                    pass
                else:
                    block[i] = r
            first = self.store.append(block)
            assert first == len(self.ids)
            for cid in new_ids:
                self.row_of[cid] = len(self.ids)
                self.ids.append(cid)
        if invalid:
            self.store.invalidate(invalid)
        if meta:
            self.meta.update(meta)

    def sync_from_dict(self, existing: Dict[str, Any]) -> None:
        """Mirror mode: make the resident rows equal to `existing` (the dict the reference would
        loop over), in its iteration order.  Fast path: same leading keys -> only rows whose value
        differs from the previous snapshot are rewritten and new keys are appended; a removed or
        reordered key rebuilds the store (the reference re-reads everything on every call anyway)."""
        keys = list(existing.keys())
        n = len(self.ids)
        prev = self._mirror_prev
        if keys[:n] != self.ids:
            self.clear()
            n, prev = 0, None
        changed = []
        if n:
            if prev is None:
                changed = keys[:n]                                            # no snapshot to compare with: rewrite
            else:
                changed = [c for c in keys[:n] if existing[c] is not prev.get(c) and existing[c] != prev.get(c)]
        if changed or len(keys) > n:
            self.upsert([(c, existing[c]) for c in changed] + [(c, existing[c]) for c in keys[n:]])
        self._mirror_prev = existing
        self.complete = False          # a LIMIT-5000, ingest-time snapshot: not the whole graph

    # -- persistence ------------------------------------------------------
    def save(self, path: str) -> None:
        """Binary sidecar next to the JSON export of src/components/graph_exporter.py:81-108: rows in the
        store dtype, the skipped mask, the id table and the per-chunk metadata."""
        import json
        if self.store is None:
            raise ValueError("nothing to save: the store has no rows yet")
        self.store.save(path, ids=self.ids, extra={"meta": json.dumps(self.meta, ensure_ascii=False, default=str)})

    @classmethod
    def load(cls, path: str, device: int = 0, initial_capacity: int = 8192) -> "ResidentChunkStore":
        """Re-hydrates the HBM store from a sidecar: bit-identical rows, same row order, same ids."""
        import json
        from .store import EmbeddingStore
        st, ids, extra = EmbeddingStore.load(path, capacity=None, device=device, with_extra=True, min_capacity=initial_capacity,
                                             max_capacity="auto")
        self = cls(st.dtype, device, initial_capacity)
        self.store, self.dim = st, st.dim
        if len(ids) != len(st):
            raise ValueError(f"sidecar holds {len(st)} rows but {len(ids)} ids")
        self.ids = list(ids)
        self.row_of = {c: i for i, c in enumerate(self.ids)}
        self.meta = json.loads(extra.get("meta", "{}"))
        return self

    def load_export(self, export: Any) -> int:
        """Hydrates from a graph export (the dict GraphExporter.export_graph writes, or its path;
        src/components/graph_exporter.py:59-68): every node labelled Chunk with an id and a list-valued
        embedding -- the filter of _get_chunk_embeddings (pre_llm_injector.py:394-409) -- in file order.
        -> number of chunk rows taken."""
        import json
        if isinstance(export, (str, bytes)) or hasattr(export, "__fspath__"):
            with open(export, "r", encoding="utf-8") as f:
                export = json.load(f)
        items, meta = [], {}
        for node in export.get("nodes", []):
            props = node.get("properties") or {}
            cid, emb = props.get("id"), props.get("embedding")
            if "Chunk" in (node.get("labels") or []) and cid and isinstance(emb, list):
                items.append((cid, emb))
                meta[cid] = {"content": props.get("content"), "time": props.get("time")}
        self.upsert(items, meta=meta)
        return len(items)

    def vectors(self, ids: Sequence[str]) -> np.ndarray:
        """The resident embeddings of `ids` as float64 [len(ids), dim] (the values the exact steps score: the originals of
        a binary64 store, the stored values otherwise); one device gather + one copy."""
        import torch
        rows = torch.as_tensor([self.row_of[c] for c in ids], dtype=torch.int64)
        src = self.store.rows_exact if getattr(self.store, "exact", False) else self.store.rows
        return src[rows.to(src.device), :self.dim].double().cpu().numpy()

    # -- queries ---------------------------------------------------------------------------------
    def topk(self, queries: Sequence[Any], k: int, min_score: float = -math.inf, score_mode: int = L.VM_SCORE_RAW,
             flags: int = 0):
        """queries: list of vectors (or Exception objects).  -> list (per query) of [(chunk_id, score)]."""
        out: List[List[Tuple[str, float]]] = [[] for _ in queries]
        if self.store is None or len(self.ids) == 0:
            return out
        good = [i for i, q in enumerate(queries)
                if not isinstance(q, Exception) and not _falsy_embedding(q) and len(q) == self.dim]
        if good:
            qa = np.asarray([queries[i] for i in good], dtype=np.float64)
            idx, score, count = self.store.topk(qa, k, min_score=min_score, score_mode=score_mode, flags=flags)
            for o, i in enumerate(good):
                out[i] = [(self.ids[int(idx[o, j])], float(score[o, j])) for j in range(int(count[o]))]
        # a query of the wrong length scores 0.0 against every row (:378-379): first k rows in store order
        for i, q in enumerate(queries):
            if isinstance(q, Exception) or i in good:
                continue
            if _falsy_embedding(q) or len(q) != self.dim:
                valid = (self.store.inv_norms[:len(self.ids)] >= 0).cpu().numpy()
                rows = np.nonzero(valid)[0][:k]
                val = 0.0 if score_mode == L.VM_SCORE_RAW else 0.5
                if val > min_score:
                    out[i] = [(self.ids[int(r)], val) for r in rows]
        return out


class ChunkSimilarityBackend:
    """S1 / S2: PreLLMInjector._calculate_batch_similarities and _cosine_similarity
    (src/components/pre_llm_injector.py:346-388) plus the cross-query merge (:235-249)."""

    def __init__(self, store: Optional[ResidentChunkStore] = None, mirror_fetch: bool = True, **store_kw):
        self.store = store if store is not None else ResidentChunkStore(**store_kw)
        #: True  = behave exactly like the reference: fetch `{id: embedding}` through the injector's own
        #:         _get_chunk_embeddings (LIMIT 5000) on every call and mirror it into HBM;
        #: False = rows arrive only through the insert hook; no Bolt fetch on the query path (row f1).
        self.mirror_fetch = mirror_fetch

    # S1 -- same signature and return type as the reference coroutine
    async def _calculate_batch_similarities(self, injector, chunk_embeddings, neo4j_handler) -> List[List[Tuple[str, float]]]:
        if self.mirror_fetch:
            existing = await injector._get_chunk_embeddings(neo4j_handler)   # reference's own fetch (:353)
            self.store.sync_from_dict(existing)
        k = injector.embedder_config.top_k_chunk_with_batch_similarity      # (:370)
        return self.store.topk(list(chunk_embeddings), k)

    # row f1: one bulk read instead of a full fetch per batch
    HYDRATE_QUERY = """
        MATCH (c:Chunk:GraphNode)
        WHERE c.graph_uuid = $graph_uuid AND c.id IS NOT NULL AND c.embedding IS NOT NULL
        RETURN c.id as chunk_id, c.embedding as embedding, c.content as content, c.time as time
    """

    async def hydrate(self, neo4j_handler, limit: Optional[int] = None, page_rows: int = 4096) -> int:
        """Fills the resident store once from Neo4j with the MATCH/WHERE/RETURN of the reference's
        _get_chunk_embeddings (src/components/pre_llm_injector.py:394-399) -- without its `LIMIT 5000`
        unless `limit` is given -- streaming the result in pages of `page_rows` rows so that no
        Python dict of the whole store is ever built.  Row order = the order Neo4j returns, exactly as
        in the reference.  Afterwards the query path needs no Bolt fetch (mirror_fetch is switched
        off); new rows arrive through on_chunks_inserted.  -> rows read."""
        query = self.HYDRATE_QUERY + (f"        LIMIT {int(limit)}\n" if limit is not None else "")
        total = 0
        self.store.clear()
        async with neo4j_handler.driver.session() as session:
            result = await session.run(query, graph_uuid=neo4j_handler.run_uuid)
            page, meta = [], {}
            async for record in result:
                chunk_id, embedding = record["chunk_id"], record["embedding"]
                if isinstance(embedding, list) and chunk_id:               # the reference's row filter (:405)
                    page.append((chunk_id, embedding))
                    get = record.get if hasattr(record, "get") else (lambda key: None)
                    meta[chunk_id] = {"content": get("content"), "time": get("time")}
                    if len(page) >= page_rows:
                        self.store.upsert(page, meta=meta)
                        total += len(page)
                        page, meta = [], {}
            if page:
                self.store.upsert(page, meta=meta)
                total += len(page)
        self.mirror_fetch = False
        self.store.complete = limit is None
        return total

    # S2
    def _cosine_similarity(self, vec1: List[float], vec2: List[float]) -> float:
        if len(vec1) != len(vec2):                                            # (:378-379)
            return 0.0
        if len(vec1) == 0:
            return 0.0
        from .store import cosine_pairs
        return float(cosine_pairs(vec1, vec2, zero_rule=0, device=self.store.device)[0])

    # cross-query merge (:235-249): max score per id, stable sort desc, [:top_k_similar_batch]
    def merge_top_similar(self, batch_similarities: List[List[Tuple[str, float]]], top_k2: int) -> List[Tuple[str, float]]:
        import ctypes as C
        import torch
        nq = len(batch_similarities)
        k = max((len(x) for x in batch_similarities), default=0)
        if nq == 0 or k == 0:
            return []
        # ids that are not resident rows (should not happen) get private negative indices
        extra: Dict[str, int] = {}
        idx = np.full((nq, k), -1, np.int64)
        sc = np.zeros((nq, k), np.float64)
        cnt = np.zeros(nq, np.int32)
        for i, lst in enumerate(batch_similarities):
            cnt[i] = len(lst)
            for j, (cid, s) in enumerate(lst):
                idx[i, j] = self.store.row_of.get(cid, extra.setdefault(cid, -2 - len(extra)))
                sc[i, j] = s
        dev = torch.device("cuda", self.store.device)
        d_idx, d_sc, d_cnt = (torch.from_numpy(a).to(dev) for a in (idx, sc, cnt))
        o_idx = torch.full((top_k2,), -1, dtype=torch.int64, device=dev)
        o_sc = torch.zeros((top_k2,), dtype=torch.float64, device=dev)
        o_cnt = torch.zeros((1,), dtype=torch.int32, device=dev)
        lib = L.load()
        L.check(lib.vm_merge_max_by_id(self.store.device, d_idx.data_ptr(), d_sc.data_ptr(), d_cnt.data_ptr(), nq, k,
                                       top_k2, o_idx.data_ptr(), o_sc.data_ptr(), o_cnt.data_ptr(),
                                       torch.cuda.current_stream(dev).cuda_stream))
        m = int(o_cnt.item())
        rev = {v: c for c, v in extra.items()}
        rows, scores = o_idx[:m].cpu().numpy(), o_sc[:m].cpu().numpy()
        return [((self.store.ids[int(r)] if r >= 0 else rev[int(r)]), float(s)) for r, s in zip(rows, scores)]

    # row f2: S1 and the cross-query merge in ONE device pass
    def similarities_and_top_similar(self, chunk_embeddings: Sequence[Any], k: int, top_k2: int):
        """-> (batch_similarities, top_similar_chunks): what _calculate_batch_similarities returns AND the merged seed list
        of pre_llm_injector.py:235-249, computed without leaving the device in between -- the per-query top-k lists feed
        `vm_merge_max_by_id` on the same stream, and both results come back after one synchronisation.  Falls back to
        the two separate calls whenever a query needs the adapter's special cases (Exception, falsy or wrong-length
        vectors) or the store is rank-routed."""
        import torch
        rs = self.store
        queries = list(chunk_embeddings)
        st = getattr(rs, "store", None)
        plain = (st is not None and hasattr(st, "topk_device") and len(getattr(rs, "ids", ())) > 0 and len(queries) > 0 and
                 all(not isinstance(q, Exception) and not _falsy_embedding(q) and len(q) == rs.dim for q in queries))
        if not plain:
            lists = rs.topk(queries, k)
            return lists, self.merge_top_similar(lists, top_k2)
        dev = st.device
        q = torch.from_numpy(np.asarray(queries, dtype=np.float64)).to(dev)
        idx, sc, cnt = st.topk_device(q, k, flags=L.VM_FLAG_ASYNC)
        o_idx = torch.full((top_k2,), -1, dtype=torch.int64, device=dev)
        o_sc = torch.zeros((top_k2,), dtype=torch.float64, device=dev)
        o_cnt = torch.zeros((1,), dtype=torch.int32, device=dev)
        L.check(L.load().vm_merge_max_by_id(dev.index, idx.data_ptr(), sc.data_ptr(), cnt.data_ptr(), len(queries), k, top_k2,
                                            o_idx.data_ptr(), o_sc.data_ptr(), o_cnt.data_ptr(),
                                            torch.cuda.current_stream(dev).cuda_stream))
        h_idx, h_sc, h_cnt, m_idx, m_sc, m_cnt = (t.cpu().numpy() for t in (idx, sc, cnt, o_idx, o_sc, o_cnt))   # first .cpu() synchronises
        lists = [[(rs.ids[int(h_idx[i, j])], float(h_sc[i, j])) for j in range(int(h_cnt[i]))] for i in range(len(queries))]
        merged = [(rs.ids[int(m_idx[j])], float(m_sc[j])) for j in range(int(m_cnt[0]))]
        return lists, merged

    # S6 -- insert hook, same `text_chunks` list Neo4jHandler._create_chunks_with_embeddings receives
    def on_chunks_inserted(self, text_chunks: List[Dict[str, Any]]) -> None:
        self.store.upsert(((c["id"], c.get("embedding")) for c in text_chunks),
                          meta={c["id"]: {"content": c.get("content"), "time": c.get("time")} for c in text_chunks})


class VectorSearchBackend:
    """S3 / S4: HybridRetriever._vector_search_chunks and _cosine_similarity
    (src/pipeline/retriever_hybrid.py:284-323, 655-664) plus post-compression (:465-514)."""

    MIN_SCORE = 0.3  # `WHERE similarity > 0.3` (:298), on the Neo4j-normalised score

    #: content/time of the returned ids when the resident store does not hold them (mirror mode keeps vectors
    #: only): the RETURN clause of the reference's Cypher (:299) restricted to the hits
    META_QUERY = """
        MATCH (c:Chunk {graph_uuid: $graph_uuid})
        WHERE c.id IN $ids
        RETURN c.id AS chunk_id, c.time AS chunk_time, c.content AS content
    """

    def __init__(self, store: ResidentChunkStore, fallback=None):
        self.store = store
        #: the reference's own bound _vector_search_chunks (set by install_retriever): used whenever the resident
        #: store is not known to hold every chunk of the graph (ResidentChunkStore.complete)
        self.fallback = fallback

    async def _fetch_meta(self, retriever, session, ids: List[str]) -> None:
        result = await session.run(self.META_QUERY, graph_uuid=retriever.neo4j_handler.run_uuid, ids=list(ids))
        async for record in result:
            self.store.meta[record["chunk_id"]] = {"content": record["content"], "time": record["chunk_time"]}

    async def _vector_search_chunks(self, retriever, session, query: str) -> List[Dict[str, Any]]:
        if not getattr(self.store, "complete", True) and self.fallback is not None:
            return await self.fallback(session, query)       # partial mirror: the reference's exhaustive Cypher scan
        try:
            query_embedding = await retriever.neo4j_handler.embedder.aembed_query(query)   # (:290)
            res = self.store.topk([query_embedding], retriever.config.top_k_chunks, min_score=self.MIN_SCORE,
                                  score_mode=L.VM_SCORE_NEO4J)[0]
            missing = [cid for cid, _ in res if "content" not in self.store.meta.get(cid, {})]
            if missing:
                await self._fetch_meta(retriever, session, missing)
            chunks = []
            for cid, score in res:
                m = self.store.meta[cid]                     # KeyError (no such chunk in Neo4j) -> [] like any failure
                chunks.append({"id": cid, "time": m.get("time"), "content": m.get("content"),
                               "score": float(score), "source": "vector"})            # (:310-316)
            return chunks
        except Exception:                                                              # (:321-323)
            return []

    def _cosine_similarity(self, vec1: List[float], vec2: List[float]) -> float:
        # zip() truncates the dot product to the shorter vector while the magnitudes use the full
        # vectors (:658-660): identical to zero-padding the shorter one (a 0.0 product leaves both the
        # naive and the Neumaier recurrence unchanged)
        n = max(len(vec1), len(vec2))
        if min(len(vec1), len(vec2)) == 0:
            return 0.0
        a = np.zeros(n, np.float64); a[:len(vec1)] = vec1
        b = np.zeros(n, np.float64); b[:len(vec2)] = vec2
        from .store import cosine_pairs
        return float(cosine_pairs(a, b, zero_rule=1, device=self.store.device)[0])

    def filter_segments(self, query_embedding: List[float], segment_embeddings: List[List[float]], threshold: float,
                        top_k: Optional[int] = None) -> List[Tuple[int, float]]:
        """Batch form of the loop at :492-504: (segment index, score) for score >= threshold (inclusive),
        original order, cut at top_k.  Segments may differ in length from the query (zip truncation, see
        _cosine_similarity): every pair is zero-padded to the longest vector of the batch."""
        if not segment_embeddings:
            return []
        from .store import cosine_pairs
        n = max([len(query_embedding)] + [len(e) for e in segment_embeddings])
        S = np.zeros((len(segment_embeddings), n), np.float64)
        for i, e in enumerate(segment_embeddings):
            S[i, :len(e)] = e
        q = np.zeros(n, np.float64)
        q[:len(query_embedding)] = query_embedding
        if n == 0:
            scores = np.zeros(len(segment_embeddings))
        else:
            scores = cosine_pairs(np.broadcast_to(q, S.shape).copy(), S, zero_rule=1, device=self.store.device)
        keep = [(i, float(sc)) for i, sc in enumerate(scores) if sc >= threshold]
        return keep if top_k is None else keep[:top_k]

    async def rerank_prefilter(self, retriever, query: str, chunks: List[Dict], keep: int) -> List[Dict]:
        """Row f4, opt-in (changes what the HTTP reranker sees, so it is OFF unless install_retriever is given
        rerank_prefilter=N): of the chunks about to be sent to the reranker (retriever_hybrid.py:516-546) only the
        `keep` most similar to the query by cosine of the RESIDENT embeddings go on, in their original order.  Chunks
        without a resident embedding cannot be scored and are always kept.  Ties -> earlier chunk."""
        if keep is None or len(chunks) <= keep:
            return chunks
        rs = self.store
        if getattr(rs, "store", None) is None or not hasattr(rs, "vectors"):     # empty, or a rank-routed store: nothing to score with
            return chunks
        known = [i for i, c in enumerate(chunks) if c.get("id") in getattr(rs, "row_of", {})
                 and float(rs.store.inv_norms[rs.row_of[c["id"]]]) >= 0.0]
        if not known or rs.dim is None:
            return chunks
        q = await retriever.neo4j_handler.embedder.aembed_query(query)
        if _falsy_embedding(q) or len(q) != rs.dim:
            return chunks
        from .store import cosine_pairs
        V = rs.vectors([chunks[i]["id"] for i in known])
        sc = cosine_pairs(np.broadcast_to(np.asarray(q, np.float64), V.shape).copy(), V, zero_rule=1, device=rs.device)
        room = max(int(keep) - (len(chunks) - len(known)), 0)
        best = sorted(range(len(known)), key=lambda t: (-sc[t], known[t]))[:room]
        chosen = {known[t] for t in best} | (set(range(len(chunks))) - set(known))
        return [c for i, c in enumerate(chunks) if i in chosen]

    async def _post_compress_chunks(self, retriever, query: str, chunks: List[Dict]) -> List[Dict]:
        """S4 caller (:465-514): same splitter, same dict shape, same order and `[:top_k]` cut -- but the segment
        embeddings are requested together (one asyncio.gather instead of a sequential HTTP loop) and scored by ONE
        batched device call instead of a Python loop per segment."""
        import asyncio
        import sys
        if not retriever.embedder or not chunks:                                       # (:467-468)
            return chunks
        try:
            query_embedding = await retriever.embedder.aembed_query(query)             # (:474)
            splitter_cls = getattr(sys.modules.get(type(retriever).__module__), "RecursiveCharacterTextSplitter")
            splitter = splitter_cls(chunk_size=256, chunk_overlap=32, separators=["\n\n", "\n", ". ", " "])   # (:478-482)
            owners, segments = [], []
            for chunk in chunks:
                for segment in splitter.split_text(chunk["content"]):                  # (:486-489)
                    owners.append(chunk)
                    segments.append(segment)
            embedded = await asyncio.gather(*(retriever.embedder.aembed_query(sg) for sg in segments),
                                            return_exceptions=True)
            ok = [i for i, e in enumerate(embedded) if not isinstance(e, BaseException)]   # a failed segment is skipped (:505-507)
            kept = self.filter_segments(query_embedding, [embedded[i] for i in ok], retriever.config.compression_threshold)
            out = [{**owners[ok[i]], "content": segments[ok[i]], "compression_score": float(sc)} for i, sc in kept]
            return out[:retriever.config.top_k]                                        # (:510)
        except Exception:                                                              # (:512-514)
            return chunks


class EmbeddingUtilsBackend:
    """EmbeddingUtils.cosine_similarity (src/utils/embedding_utils.py:29-39) -- zip-truncated
    dot product, magnitudes through `** 0.5`, 0.0 when either magnitude is zero."""

    def __init__(self, device: int = 0):
        self.device = device

    def cosine_similarity(self, vec1: List[float], vec2: List[float]) -> float:
        n = max(len(vec1), len(vec2))
        if min(len(vec1), len(vec2)) == 0:
            return 0.0
        a = np.zeros(n, np.float64); a[:len(vec1)] = vec1
        b = np.zeros(n, np.float64); b[:len(vec2)] = vec2
        from .store import cosine_pairs
        return float(cosine_pairs(a, b, zero_rule=2, device=self.device)[0])


class PruneBackend:
    """S5: Graph._are_same_context / _get_representative_relation (src/pipeline/prune.py:56-79)."""

    def __init__(self, device: int = 0):
        self.device = device

    def _are_same_context(self, graph, relation_sentences, threshold: float = 0.8) -> bool:
        if len(relation_sentences) <= 1:                                   # (:73-74)
            return False
        emb = np.asarray(graph.embedding_model.encode(relation_sentences), dtype=np.float32)
        from .dedup import pairs_above
        import torch
        i, j, s = pairs_above(torch.from_numpy(emb).to(f"cuda:{self.device}"), float(threshold))
        return bool(len(i) > 0)                                           # np.any(S > threshold) (:79)

    def _get_representative_relation(self, graph, relation_sentences) -> int:
        emb = np.asarray(graph.embedding_model.encode(relation_sentences), dtype=np.float32)
        centroid = np.mean(emb, axis=0)                                    # (:61)
        from .store import EmbeddingStore
        st = EmbeddingStore(emb.shape[1], len(emb), "f32", self.device)
        try:
            st.append(emb)
            idx, score, count = st.topk(centroid[None, :].astype(np.float32), 1)
            return int(idx[0, 0]) if count[0] else 0                       # argmax, first maximal index (:64)
        finally:
            st.close()


# ---- binding onto unmodified reference objects ---------------------------------------------------
def install_injector(injector, backend: Optional[ChunkSimilarityBackend] = None, **kw) -> ChunkSimilarityBackend:
    """Rebinds S1/S2 on a reference PreLLMInjector instance; everything else is untouched."""
    backend = backend or ChunkSimilarityBackend(**kw)

    async def _calc(self, chunk_embeddings, neo4j_handler):
        return await backend._calculate_batch_similarities(self, chunk_embeddings, neo4j_handler)

    def _cos(self, vec1, vec2):
        return backend._cosine_similarity(vec1, vec2)

    injector._calculate_batch_similarities = types.MethodType(_calc, injector)
    injector._cosine_similarity = types.MethodType(_cos, injector)
    return backend


def install_retriever(retriever, store: ResidentChunkStore, rerank_prefilter: Optional[int] = None) -> VectorSearchBackend:
    """Rebinds S3 (_vector_search_chunks), S4 (_cosine_similarity) and S4's caller (_post_compress_chunks) on a
    reference HybridRetriever instance.  The instance's own _vector_search_chunks is kept as the fallback for a
    store that does not hold the whole graph (mirror mode).  rerank_prefilter=N (opt-in, row f4): _rerank_chunks only
    sends the N chunks closest to the query (cosine of the resident embeddings) to the HTTP reranker."""
    backend = VectorSearchBackend(store, fallback=getattr(retriever, "_vector_search_chunks", None))

    async def _vs(self, session, query):
        return await backend._vector_search_chunks(self, session, query)

    async def _pc(self, query, chunks):
        return await backend._post_compress_chunks(self, query, chunks)

    retriever._vector_search_chunks = types.MethodType(_vs, retriever)
    retriever._post_compress_chunks = types.MethodType(_pc, retriever)
    retriever._cosine_similarity = backend._cosine_similarity      # static in the reference: called with (vec1, vec2)
    inner_rerank = getattr(retriever, "_rerank_chunks", None)
    if rerank_prefilter is not None and inner_rerank is not None:
        async def _rr(self, query, chunks, raise_on_failure=False):
            if getattr(self.config, "use_reranker", False) and chunks:      # the reference returns early otherwise (:518)
                chunks = await backend.rerank_prefilter(self, query, chunks, rerank_prefilter)
            return await inner_rerank(query, chunks, raise_on_failure)

        retriever._rerank_chunks = types.MethodType(_rr, retriever)
    return backend


def install_embedding_utils(target, backend: Optional[EmbeddingUtilsBackend] = None, **kw) -> EmbeddingUtilsBackend:
    """Rebinds cosine_similarity on the reference's EmbeddingUtils class (or an instance of it)."""
    backend = backend or EmbeddingUtilsBackend(**kw)
    fn = backend.cosine_similarity
    if isinstance(target, type):
        target.cosine_similarity = staticmethod(fn)
    else:
        target.cosine_similarity = fn
    return backend


def install_prune(graph, backend: Optional[PruneBackend] = None) -> PruneBackend:
    backend = backend or PruneBackend()
    graph._are_same_context = types.MethodType(lambda self, s, threshold=0.8: backend._are_same_context(self, s, threshold), graph)
    graph._get_representative_relation = types.MethodType(lambda self, s: backend._get_representative_relation(self, s), graph)
    return backend
