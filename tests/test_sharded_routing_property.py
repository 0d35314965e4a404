"""Property test of the rank-routed streaming store's host logic: G ranks simulated in one process (no
process group), the device store replaced by the oracle-backed double, the exchange by merge_lists_host.
For random insert sequences -- new ids, repeated ids (upserts), falsy embeddings, ids repeated inside one
batch -- and random queries, the merged answer equals the oracle over the global store in first-seen order."""
import numpy as np
from hypothesis import given, settings, strategies as st

from doubles import OracleBackedStore
from oracle import oracle


def _simulate(world, batches, queries, k, monkeypatch_target):
    from vidmem_b200 import sharded
    stores = [sharded.ShardedChunkStore("f32", device=0, rank=r, world=world) for r in range(world)]
    for items in batches:
        for s in stores:
            s.upsert(items)
    lists = [s._local_lists(queries, k, -np.inf, 0, 0) for s in stores]
    merged = sharded.merge_lists_host(lists, k)
    named = [s._named(*merged) for s in stores]
    return stores, named


@st.composite
def scenario(draw):
    d = draw(st.integers(2, 6))
    world = draw(st.integers(1, 4))
    n_batches = draw(st.integers(1, 6))
    pool = draw(st.integers(1, 12))                                   # id pool: small -> many upserts
    vec = st.lists(st.integers(-4, 4).map(float), min_size=d, max_size=d)
    batches = []
    for _ in range(n_batches):
        size = draw(st.integers(1, 5))
        items = []
        for _ in range(size):
            cid = f"c{draw(st.integers(0, pool - 1))}"
            emb = draw(st.one_of(vec, st.just([]), st.none()))
            items.append((cid, emb))
        batches.append(items)
    queries = draw(st.lists(vec, min_size=1, max_size=3))
    k = draw(st.integers(1, 4))
    return d, world, batches, queries, k


@settings(max_examples=60, deadline=None)
@given(scenario())
def test_routed_store_equals_oracle_over_global_order(sc):
    import vidmem_b200.store as vstore
    d, world, batches, queries, k = sc
    saved = vstore.EmbeddingStore
    vstore.EmbeddingStore = OracleBackedStore
    try:
        stores, named = _simulate(world, batches, queries, k, None)
    finally:
        vstore.EmbeddingStore = saved
    # reference semantics: dict insertion order, last write wins, falsy embeddings skipped (:363)
    order, emb = [], {}
    for items in batches:
        for cid, e in items:
            if cid not in emb:
                order.append(cid)
            emb[cid] = e
    assert all(s.ids == order for s in stores)
    assert sum(stores[0].load) == len(order) and all(s.load == stores[0].load for s in stores)
    # every id lives on exactly the rank the shared table says
    for r, s in enumerate(stores):
        assert s.local.ids == [c for c, o in zip(order, s.owner) if o == r]
    if not any(emb[c] for c in order):
        assert all(lst == [] for lst in named[0])                      # nothing embeddable yet
        return
    X = np.array([emb[c] if emb[c] else [0.0] * d for c in order], np.float64)
    ok = np.array([1 if emb[c] else 0 for c in order], np.uint8)
    want = [[(order[r], s) for r, s in lst] for lst in oracle.batch_similarities(np.array(queries, np.float64), X, k, row_ok=ok)]
    for got in named:
        assert got == want
