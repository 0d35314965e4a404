"""N>1 host logic on CPU: shard bounds, and a world_size-2 gloo run in which every rank scores its
row shard (with the CPU oracle standing in for the device scan), the per-rank lists are gathered and
merged exactly like vm_topk_sharded's device merge (score desc, global row asc)."""
import os
import socket

import numpy as np
import pytest

from oracle import oracle, synth


def _sharded():
    import vidmem_b200
    from vidmem_b200 import sharded
    return sharded


def test_shard_bounds_partition_rows():
    sh = _sharded()
    for n in (0, 1, 7, 100, 1000003):
        for g in (1, 2, 3, 8):
            b = sh.shard_bounds(n, g)
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(g - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
            for row in (0, n // 3, n - 1):
                if 0 <= row < n:
                    r = sh.owner_of(row, n, g)
                    assert b[r][0] <= row < b[r][1]


def test_merge_lists_host_ties_and_short_lists():
    sh = _sharded()
    a = (np.array([[5, 9, -1]]), np.array([[0.9, 0.5, 0.0]]), np.array([2], np.int32))
    b = (np.array([[12, 3, 40]]), np.array([[0.9, 0.9, 0.1]]), np.array([3], np.int32))
    i, s, c = sh.merge_lists_host([a, b], 4)
    assert list(i[0]) == [3, 5, 12, 9] and list(s[0]) == [0.9, 0.9, 0.9, 0.5] and c[0] == 4
    i, s, c = sh.merge_lists_host([a, b], 10)
    assert c[0] == 5 and i[0, 5] == -1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, d, nq, k, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vidmem_b200 import sharded
    lo, hi = sharded.shard_bounds(n, world)[rank]
    X = synth.synth_rows(7, lo, hi - lo, d)
    X_full = synth.synth_rows(7, 0, n, d)
    X_full[n - 1] = X_full[0]                      # a tie that straddles the two shards
    X = X_full[lo:hi]
    Q = synth.synth_queries(8, nq, d, 7, n)
    Q[0] = X_full[0]
    local = oracle.batch_similarities(Q, X, k)     # stand-in for the device scan + exact rescoring
    idx = np.full((nq, k), -1, np.int64); sc = np.zeros((nq, k)); cnt = np.zeros(nq, np.int32)
    for qi, lst in enumerate(local):
        cnt[qi] = len(lst)
        for j, (r, s) in enumerate(lst):
            idx[qi, j], sc[qi, j] = r + lo, s    # global row indices, as vm_topk_sharded emits
    gathered = [None] * world
    dist.all_gather_object(gathered, (idx, sc, cnt))
    mi, ms, mc = sharded.merge_lists_host(gathered, k)
    ref = oracle.batch_similarities(Q, X_full, k)
    ok = all(list(mi[q, :mc[q]]) == [r for r, _ in ref[q]] and list(ms[q, :mc[q]]) == [s for _, s in ref[q]]
             for q in range(nq))
    ok = ok and list(mi[0, :2]) == [0, n - 1]
    open(os.path.join(out_dir, f"rank{rank}.ok" if ok else f"rank{rank}.bad"), "w").write("x")
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo(tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, 2001, 64, 5, 10, str(tmp_path)), nprocs=2, join=True)
    assert sorted(os.listdir(tmp_path)) == ["rank0.ok", "rank1.ok"]


def _stream_scenario(d=48):
    """Insert batches (ids, rows) + queries shared by the CPU (gloo) and the 2-GPU test."""
    X = synth.synth_rows(31, 0, 200, d)
    batches, row = [], 0
    for b, size in enumerate((5, 40, 3, 17, 64, 1, 30)):
        items = [(f"u_{b}_{i}", [float(v) for v in X[row + i]]) for i in range(size)]
        row += size
        batches.append(items)
    batches[3][2] = ("u_3_2", [])                                         # falsy embedding: a skipped row (:363)
    batches[5][0] = ("u_5_0", batches[0][1][1])                           # duplicate of an earlier row -> a cross-rank tie
    batches.append([("u_1_7", [float(v) for v in X[190]]), ("u_7_0", [float(v) for v in X[191]])])  # upsert + new
    Q = synth.synth_queries(32, 6, d, 31, 160)
    Q[0] = X[1]
    Q[1] = X[190]
    return batches, [[float(v) for v in q] for q in Q]


def _stream_expected(batches, queries, k):
    order, emb = [], {}
    for items in batches:
        for cid, e in items:
            if cid not in emb:
                order.append(cid)
            emb[cid] = e
    d = len(queries[0])
    Xg = np.array([emb[c] if emb[c] else [0.0] * d for c in order], np.float64).astype(np.float32).astype(np.float64)
    ok = np.array([1 if emb[c] else 0 for c in order], np.uint8)
    ref = oracle.batch_similarities(np.array(queries), Xg, k, row_ok=ok)
    return [[(order[r], s) for r, s in lst] for lst in ref], order


def _stream_worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import vidmem_b200.store as vstore
    from vidmem_b200 import sharded
    from doubles import OracleBackedStore
    vstore.EmbeddingStore = OracleBackedStore                            # no GPU here: the oracle answers for the shard

    def host_merge(self, gathered, nq, k):                               # stands in for vm_merge_topk_lists
        g = gathered.numpy()
        lists = [(g[r, :nq * k].reshape(nq, k), g[r, nq * k:2 * nq * k].copy().view(np.float64).reshape(nq, k),
                  g[r, 2 * nq * k:].astype(np.int32)) for r in range(g.shape[0])]
        return sharded.merge_lists_host(lists, k)

    sharded.ShardedChunkStore._merge = host_merge
    batches, queries = _stream_scenario()
    st = sharded.ShardedChunkStore("f32", device=0)
    for items in batches:
        st.upsert(items)
    k = 4
    got = st.topk(queries, k)
    want, order = _stream_expected(batches, queries, k)
    ok = got == want and st.ids == order
    ok = ok and [c for c, _ in got[0][:2]] == ["u_0_1", "u_5_0"]          # the tie resolves to the earlier global row
    ok = ok and got[1][0][0] == "u_1_7"                                   # the overwritten row answers with its new vector
    ok = ok and sorted(set(st.owner)) == [0, 1] and abs(st.load[0] - st.load[1]) <= 64
    ok = ok and len(st.local) == st.load[rank]
    # the S1 / S6 adapters take the sharded store unchanged (same surface as ResidentChunkStore)
    import asyncio
    import types
    from vidmem_b200 import adapters
    st2 = sharded.ShardedChunkStore("f32", device=0)
    backend = adapters.ChunkSimilarityBackend(store=st2, mirror_fetch=False)
    assert backend.store is st2                                           # an empty store must not be replaced
    for items in batches:
        backend.on_chunks_inserted([{"id": cid, "content": f"text of {cid}", "embedding": emb} for cid, emb in items])
    inj = types.SimpleNamespace(embedder_config=types.SimpleNamespace(top_k_chunk_with_batch_similarity=k))
    got2 = asyncio.run(backend._calculate_batch_similarities(inj, queries + [RuntimeError("embed failed")], None))
    ok = ok and got2[:-1] == want and got2[-1] == [] and st2.meta["u_1_7"]["content"] == "text of u_1_7"
    open(os.path.join(out_dir, f"rank{rank}.ok" if ok else f"rank{rank}.bad"), "w").write("x")
    dist.barrier()
    dist.destroy_process_group()


def test_streaming_inserts_world_size_2_gloo(tmp_path):
    import torch.multiprocessing as mp
    mp.spawn(_stream_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert sorted(os.listdir(tmp_path)) == ["rank0.ok", "rank1.ok"]


def _dedup_worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vidmem_b200 import dedup
    E = synth.synth_rows(44, 0, 900, 64, dup_period=5)
    oi, oj, os_ = oracle.pairs_above(E, 0.9)

    def part_of_oracle(x, threshold, cap=1 << 20, part=0, nparts=1):
        # stands in for vm_pairs_above(part, nparts): this rank's share of the upper-triangular 128x128 tile grid
        mine = ((oi // 128) * 7 + (oj // 128)) % nparts == part
        return oi[mine], oj[mine], os_[mine].astype(np.float32)

    dedup.pairs_above = part_of_oracle
    gi, gj, gs = dedup.pairs_above_sharded(torch.from_numpy(E), 0.9)
    ok = len(oi) > 20 and np.array_equal(gi, oi) and np.array_equal(gj, oj) and np.array_equal(gs, os_.astype(np.float32))
    shares = [int((((oi // 128) * 7 + (oj // 128)) % world == r).sum()) for r in range(world)]
    ok = ok and min(shares) > 0                                           # both ranks really contributed
    open(os.path.join(out_dir, f"rank{rank}.ok" if ok else f"rank{rank}.bad"), "w").write("x")
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_dedup_concatenation_world_size_2_gloo(tmp_path):
    """Host side of the multi-GPU all-pairs path: per-rank hit lists of different lengths are exchanged (counts,
    then a padded all-gather), concatenated and sorted by (i, j) -- identical on every rank."""
    import torch.multiprocessing as mp
    mp.spawn(_dedup_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert sorted(os.listdir(tmp_path)) == ["rank0.ok", "rank1.ok"]
