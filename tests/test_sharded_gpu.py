"""2-GPU test of vm_topk_sharded: row-sharded store, one ncclAllGather, device merge."""
import os
import socket

import numpy as np
import pytest

from oracle import oracle, synth

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import vidmem_b200 as vm
    from vidmem_b200.sharded import Communicator, shard_bounds
    comm = Communicator.from_torch_distributed(rank)
    ok = True
    for dtype, n, d, nq, k in (("f32", 50001, 384, 64, 10), ("bf16", 20000, 128, 5, 3)):
        lo, hi = shard_bounds(n, world)[rank]
        Xf = synth.synth_rows(17, 0, n, d)
        Xf[n - 1] = Xf[2]
        Q = synth.synth_queries(18, nq, d, 17, n)
        Q[0] = Xf[2]
        st = vm.EmbeddingStore(d, hi - lo, dtype, device=rank)
        st.append(Xf[lo:hi])
        idx, score, count = st.topk(Q, k, comm=comm, row_offset=lo, sum_mode=vm.VM_SUM_NEUMAIER)
        ref = oracle.batch_similarities(Q, Xf, k)
        for qi, lst in enumerate(ref):
            ok = ok and count[qi] == len(lst) and list(idx[qi]) == [r for r, _ in lst] and list(score[qi]) == [s for _, s in lst]
        ok = ok and list(idx[0, :2]) == [2, n - 1]
        qd = torch.from_numpy(Q).cuda()
        di, ds, dc = st.topk_device(qd, k, comm=comm, row_offset=lo, sum_mode=vm.VM_SUM_NEUMAIER, flags=vm.VM_FLAG_ASYNC)
        torch.cuda.synchronize()
        ok = ok and np.array_equal(di.cpu().numpy(), idx) and np.array_equal(ds.cpu().numpy(), score)
        st.close()
    # peer-memory exchange (symmetric memory + one pull-merge kernel) == NCCL all-gather path, across
    # several consecutive batches so both slots are reused
    n, d, nq, k = 30011, 256, 150, 10
    lo, hi = shard_bounds(n, world)[rank]
    Xf = synth.synth_rows(19, 0, n, d)
    Q = synth.synth_queries(20, nq, d, 19, n)
    st = vm.EmbeddingStore(d, hi - lo, "f32", device=rank)
    st.append(Xf[lo:hi])
    base = st.topk(Q, k, comm=comm, row_offset=lo, sum_mode=vm.VM_SUM_NEUMAIER)
    ref = oracle.batch_similarities(Q, Xf, k)
    ok = ok and all(list(base[0][qi]) == [r for r, _ in ref[qi]] for qi in range(nq))
    enabled = comm.enable_peer_exchange()
    peer_state = "p2p" if enabled else "nccl-only:" + getattr(comm, "peer_exchange_error", "?")
    if enabled:
        for rep in range(5):
            got = st.topk(Q, k, comm=comm, row_offset=lo, sum_mode=vm.VM_SUM_NEUMAIER)
            ok = ok and all(np.array_equal(a, b) for a, b in zip(base, got))
        qd = torch.from_numpy(Q).cuda()
        for rep in range(4):
            di, ds, dc = st.topk_device(qd, k, comm=comm, row_offset=lo, sum_mode=vm.VM_SUM_NEUMAIER, flags=vm.VM_FLAG_ASYNC)
        torch.cuda.synchronize()
        ok = ok and np.array_equal(di.cpu().numpy(), base[0]) and np.array_equal(ds.cpu().numpy(), base[1])
    open(os.path.join(out_dir, f"rank{rank}.{peer_state[:3]}"), "w").write(peer_state)
    st.close()
    # all-pairs dedup split over the ranks == the single-GPU pair set
    from vidmem_b200 import dedup
    E = synth.synth_rows(44, 0, 5000, 256, dup_period=6)
    x = torch.from_numpy(E).cuda().to(torch.bfloat16)
    gi, gj, gs = dedup.pairs_above_sharded(x, 0.9)
    oi, oj, _ = oracle.pairs_above(E, 0.9)
    ok = ok and len(oi) > 0 and list(zip(gi.tolist(), gj.tolist())) == list(zip(oi.tolist(), oj.tolist()))
    # the same through the C ABI alone (vm_pairs_above_sharded): only rank 0 holds the rows, ONE ncclBroadcast
    # replicates them, ncclAllGather of counts + lists, device concatenation
    x2 = x.clone() if rank == 0 else torch.zeros_like(x)
    ci, cj, cs = dedup.pairs_above_sharded(x2, 0.9, comm=comm, root=0)
    ok = ok and list(zip(ci.tolist(), cj.tolist())) == list(zip(oi.tolist(), oj.tolist())) and torch.equal(x2, x)
    ok = ok and np.array_equal(cs, gs)
    ai, aj, _ = dedup.pairs_above_sharded(x, 0.9, comm=comm, cap=4096)          # already replicated (root = -1)
    ok = ok and list(zip(ai.tolist(), aj.tolist())) == list(zip(oi.tolist(), oj.tolist()))
    # streaming inserts routed over the ranks + one all-gather + vm_merge_topk_lists
    from vidmem_b200.sharded import ShardedChunkStore
    from test_sharded_cpu import _stream_scenario, _stream_expected
    batches, queries = _stream_scenario()
    for dt in ("f32", "bf16"):
        sst = ShardedChunkStore(dt, device=rank)
        for items in batches:
            sst.upsert(items)
        got = sst.topk(queries, 4)
        want, order = _stream_expected(batches, queries, 4)                 # values are exact in bf16 too (ints / 128)
        ok = ok and got == want and sst.ids == order and len(sst.local) == sst.load[rank]
    open(os.path.join(out_dir, f"rank{rank}.ok" if ok else f"rank{rank}.bad"), "w").write("x")
    dist.barrier()
    comm.close()
    dist.destroy_process_group()


def test_two_gpu_sharded_topk(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    files = sorted(os.listdir(tmp_path))
    assert "rank0.ok" in files and "rank1.ok" in files, files
    assert "rank0.p2p" in files and "rank1.p2p" in files, [open(os.path.join(tmp_path, f)).read() for f in files]
