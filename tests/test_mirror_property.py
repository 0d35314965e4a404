"""Property test of mirror mode (ResidentChunkStore.sync_from_dict): after every snapshot of the dict the
reference would iterate -- grown at the end, values changed in place, keys removed or reordered -- the store
answers exactly like the reference formula over that snapshot."""
import numpy as np
from hypothesis import given, settings, strategies as st

from doubles import OracleBackedStore
from oracle import oracle


@st.composite
def snapshots(draw):
    d = draw(st.integers(2, 5))
    vec = st.lists(st.integers(-3, 3).map(float), min_size=d, max_size=d)
    cur = {}
    snaps = []
    for _ in range(draw(st.integers(1, 5))):
        op = draw(st.sampled_from(["grow", "grow", "change", "remove", "reorder"]))
        cur = dict(cur)
        if op == "grow" or not cur:
            for _ in range(draw(st.integers(1, 4))):
                cur[f"c{draw(st.integers(0, 15))}"] = draw(st.one_of(vec, st.just([])))
        elif op == "change":
            key = draw(st.sampled_from(sorted(cur)))
            cur[key] = draw(st.one_of(vec, st.just([])))
        elif op == "remove":
            cur.pop(draw(st.sampled_from(sorted(cur))))
        else:
            keys = draw(st.permutations(sorted(cur)))
            cur = {k: cur[k] for k in keys}
        snaps.append({k: list(v) for k, v in cur.items()})            # fresh list objects, as a Bolt fetch would return
    queries = draw(st.lists(vec, min_size=1, max_size=2))
    return d, snaps, queries, draw(st.integers(1, 3))


@settings(max_examples=80, deadline=None)
@given(snapshots())
def test_mirror_mode_tracks_every_snapshot(sc):
    import vidmem_b200.store as vstore
    from vidmem_b200 import adapters
    d, snaps, queries, k = sc
    saved = vstore.EmbeddingStore
    vstore.EmbeddingStore = OracleBackedStore
    try:
        from vidmem_b200 import sharded
        store = adapters.ResidentChunkStore()
        routed = sharded.ShardedChunkStore("f32", device=0, rank=0, world=1)      # same contract on the routed store
        for snap in snaps:
            store.sync_from_dict(snap)
            got = store.topk(queries, k)
            routed.sync_from_dict(snap)
            assert routed.ids == store.ids
            assert routed._named(*sharded.merge_lists_host([routed._local_lists(queries, k, -np.inf, 0, 0)], k)) == got
            keys = list(snap)
            if not any(snap[c] for c in keys):
                assert all(lst == [] for lst in got)
                continue
            X = np.array([snap[c] if snap[c] else [0.0] * d for c in keys], np.float64)
            ok = np.array([1 if snap[c] else 0 for c in keys], np.uint8)
            want = [[(keys[r], s) for r, s in lst] for lst in oracle.batch_similarities(np.array(queries, np.float64), X, k, row_ok=ok)]
            assert got == want
    finally:
        vstore.EmbeddingStore = saved
