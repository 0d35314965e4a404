#!/usr/bin/env python
"""bench.py -- headline benchmark of the embedding-similarity hot path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c2|c3|c5]

A "step" is one pass of the hot path over one batch: 64 queries, top-10, scored against every
row of the HBM-resident synthetic store (SURVEY.md 8d generator).
  N = 1 : config C2 -- 1 000 000 x 384 fp32 store on one B200 (BASELINE.json configs[1]).
  N > 1 : config C3 -- 100 000 000 x 384 bf16 store row-sharded over N GPUs (100M/N rows each),
          local scan + exact rescoring, ONE ncclAllGather of the per-rank lists, device merge
          (strong scaling: the store is fixed, rows per GPU shrink with N).
One JSON line on stdout (rank 0).  `value` = queries/s with queries and outputs resident in
HBM; `e2e` = the same metric through the public host-buffer API (pinned host queries in, host
results out, copies inside the timed region).  `--impl reference` times the reference's CPU
algorithm (oracle port, all host threads) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "queries_per_sec_top10_384d"
UNIT = "queries/s"
CONFIGS = {
    # name: (rows_total, dim, store dtype, queries, k, store seed, query seed)
    "c1": (5_000, 384, "f32", 30, 10, 1, 1001),      # the reference's own size (LIMIT 5000 store, 30 benchmark queries)
    "c2": (1_000_000, 384, "f32", 64, 10, 2, 2002),
    "c3": (100_000_000, 384, "bf16", 64, 10, 3, 3003),
    "c5": (10_000_000, 384, "f32", 1, 10, 5, 5005),
}
# all-pairs dedup: rows, dim, dtype, threshold, seed, planted-duplicate period
C4 = (1_000_000, 768, "bf16", 0.9, 4, 100)


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clocks and throttle reasons sampled DURING the timed region (in-process NVML polling,
    ~2 ms period: a subprocess `nvidia-smi -lms` starts too slowly for a region of tens of ms)."""

    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, gpu_index: int):
        self.gpu, self.samples, self.mask, self.smax, self.err = gpu_index, [], 0, None, None
        self._stop = threading.Event()
        self.t = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.gpu
            if vis:
                try:
                    idx = int(vis.split(",")[self.gpu])
                except Exception:
                    idx = self.gpu
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.nv = pynvml
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # pragma: no cover
            self.err = repr(e)
            return
        self.t = threading.Thread(target=self._poll, daemon=True)
        self.t.start()

    def _poll(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    self.mask |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    self.mask |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            except Exception as e:
                self.err = repr(e)
                return
            time.sleep(0.002)

    def stop(self):
        self._stop.set()
        if self.t:
            self.t.join(timeout=1)
        reasons = sorted(n for bit, n in self.REASONS.items() if self.mask & bit)
        if self.err and not self.samples:
            reasons = ["nvml unavailable: " + self.err]
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.smax,
                "reasons": reasons, "samples": len(self.samples)}


def _ncu_traffic(cfg: str, scan_kernel: int, rows_local: int):
    """dram__bytes_read.sum + dram__bytes_write.sum of the scan kernel from the committed ncu --set full
    capture of this workload (profiles/), per launch; None when no capture exists for this shard size."""
    p = os.path.join(ROOT, "profiles", f"r1_scan_tc_{cfg}_summary.json")
    if scan_kernel == 2 and os.path.exists(p):
        try:
            d = json.load(open(p))
            if int(d.get("rows", rows_local)) != int(rows_local):
                return None
            return d["dram_bytes_read"] + d["dram_bytes_write"]
        except Exception:
            return None
    return None


# ---------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port (C restatement of the reference's Python loop)
# ---------------------------------------------------------------------------------------------
def cpu_reference_leg(cfg_name: str, steps: int, warmup: int, sample_rows: int):
    """Times the reference algorithm (pre_llm_injector.py:346-388 restated in C, OpenMP over
    queries) on `sample_rows` rows of the same synthetic store; returns (queries/s at the FULL
    store size by linear extrapolation in rows -- the algorithm is a flat loop over rows --,
    ms per sampled step, description)."""
    import numpy as np
    from oracle import oracle, synth
    rows_total, dim, _dt, nq, k, sseed, qseed = CONFIGS[cfg_name]
    sample_rows = min(sample_rows, rows_total)
    X = synth.synth_rows(sseed, 0, sample_rows, dim).astype(np.float64)
    Q = synth.synth_queries(qseed, nq, dim, sseed, rows_total).astype(np.float64)
    oracle.lib()
    for _ in range(warmup):
        oracle.batch_similarities(Q[:8], X[:2000], k)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        oracle.batch_similarities(Q, X, k)
        times.append(time.perf_counter() - t0)
    per_step = sum(times) / len(times)
    full = per_step * (rows_total / sample_rows)
    return nq / full, per_step * 1e3, (f"{nq} queries x {sample_rows} rows x {dim} (of {rows_total} rows), "
                                       f"time scaled linearly in rows to the full store")


def cpu_blas_leg(cfg_name: str, sample_rows: int):
    """The reference's own vectorised idiom (src/pipeline/prune.py:62,76): sklearn cosine_similarity (BLAS
    sgemm over every host core, store re-normalised on every call) + argpartition/sort for the top-k, on the
    same bounded sample.  Context only: float32, not the bit-exact reference order."""
    import numpy as np
    from sklearn.metrics.pairwise import cosine_similarity
    from oracle import synth
    rows_total, dim, _dt, nq, k, sseed, qseed = CONFIGS[cfg_name]
    sample_rows = min(sample_rows, rows_total)
    X = synth.synth_rows(sseed, 0, sample_rows, dim)
    Q = synth.synth_queries(qseed, nq, dim, sseed, rows_total)
    cosine_similarity(Q[:2], X[:1000])
    t0 = time.perf_counter()
    S = cosine_similarity(Q, X)
    part = np.argpartition(-S, k - 1, axis=1)[:, :k]
    np.take_along_axis(S, part, 1).sort(axis=1)
    dt = time.perf_counter() - t0
    return nq / (dt * rows_total / sample_rows)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun pins OMP_NUM_THREADS=1 in every worker; the reference arm is meant to use every host core
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    cfg = args.config or ("c2" if args.gpus == 1 else "c3")
    cores = os.cpu_count() or 1
    sample = args.cpu_sample_rows or 100000
    steps = max(1, min(args.steps, 5))
    qps, ms, desc = cpu_reference_leg(cfg, steps, min(args.warmup, 1), sample)
    rows_total, dim, dt, nq, k, _, _ = CONFIGS[cfg]
    line = {"impl": "reference", "metric": METRIC, "value": qps, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": min(args.warmup, 1), "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak" if args.gpus == 1 else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{cfg}: {rows_total}x{dim} {dt} store, {nq}-query batch, top-{k} cosine",
                       "timed_sample": desc},
            "cpu_baseline": {"value": qps, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
            "e2e": {"value": qps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def make_queries(store, n_rows: int, nq: int, dim: int, seed: int, dev, world: int = 1):
    """Synthetic query batch for the timed arm (no oracle code on this path): even queries are copies of a
    resident row with every 4th column redrawn, odd queries are independent draws; all values are multiples
    of 1/128 like the store rows.  Under torchrun rank 0's batch is broadcast so every rank scores the same
    queries.  -> float32 CPU tensor [nq, dim]."""
    import torch
    g = torch.Generator(device="cpu").manual_seed(seed)
    ridx = torch.randint(0, max(n_rows, 1), (nq,), generator=g)
    base = store.rows[ridx.to(dev), :dim].float()
    noise = (torch.randint(-127, 128, (nq, dim), generator=g).float() / 128.0).to(dev)
    redraw = (torch.arange(dim, device=dev) % 4 == 0)[None, :] | (torch.arange(nq, device=dev) % 2 == 1)[:, None]
    q = torch.where(redraw, noise, base).contiguous()
    if world > 1:
        import torch.distributed as dist
        dist.broadcast(q, src=0)
    return q.cpu()


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import vidmem_b200 as vm
    from vidmem_b200.store import EmbeddingStore

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    comm = None
    peer_exchange = False
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        from vidmem_b200.sharded import Communicator
        comm = Communicator.from_torch_distributed(local_rank)
        peer_exchange = args.peer_exchange and comm.enable_peer_exchange()

    cfg = args.config or ("c2" if world == 1 else "c3")
    rows_total, dim, dt, nq, k, sseed, qseed = CONFIGS[cfg]
    if args.rows:
        rows_total = args.rows
    row_lo = rows_total * rank // world
    row_hi = rows_total * (rank + 1) // world
    n_local = row_hi - row_lo
    es = 4 if dt == "f32" else 2

    store = EmbeddingStore(dim, n_local, dt, device=local_rank)
    store.synth_fill(sseed, n_local, row0=row_lo)
    store.set_size(n_local)
    torch.cuda.synchronize()
    q_pinned = make_queries(store, n_local, nq, dim, qseed, dev, world).pin_memory()
    Q = q_pinned.numpy()
    q_dev = q_pinned.to(dev)
    out = (torch.empty((nq, k), dtype=torch.int64, device=dev), torch.empty((nq, k), dtype=torch.float64, device=dev),
           torch.empty((nq,), dtype=torch.int32, device=dev))
    flags_dev = vm.VM_FLAG_ASYNC | vm.VM_FLAG_TIMING | args.flags

    def step_device():
        store.topk_device(q_dev, k, out=out, flags=flags_dev, comm=comm, row_offset=row_lo)

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    # ---- leg 1: inputs resident in HBM ------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    store.avg_scan_ms()  # drop the warm-up steps' event pairs: only the timed region is averaged below
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    scan_ms = []
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fits_l2 = n_local * dim * es <= 126e6
    barrier()
    if not fits_l2:
        ev0.record()
        for _ in range(args.steps):
            step_device()
        ev1.record()
        barrier()
        dev_ms = ev0.elapsed_time(ev1)
    else:
        # store fits the 126 MB L2: flush it (write a 256 MB buffer) before every step and time the steps one by one
        flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
        dev_ms = 0.0
        for _ in range(args.steps):
            flush.zero_()
            ev0.record()
            step_device()
            ev1.record()
            ev1.synchronize()
            dev_ms += ev0.elapsed_time(ev1)
        barrier()
        del flush
    launches = int(store.last_stats.scan_launches) * args.steps
    stats = store.last_stats
    # scan-kernel time: every timed step above carried its own CUDA event pair on the launching stream
    # (VM_FLAG_TIMING); read them back now -- the average over the last <= 64 steps of the timed region itself
    scan_avg_live, scan_calls = store.avg_scan_ms()
    scan_ms = [scan_avg_live]
    barrier()

    # ---- leg 2: end to end through the host-buffer API ----------------------------------------
    def step_host():
        return store.topk(q_pinned.numpy(), k, comm=comm, row_offset=row_lo, flags=args.flags)

    for _ in range(3):
        res = step_host()
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        res = step_host()
    e1.record()
    barrier()
    e2e_wall_ms = (time.perf_counter() - t0) * 1e3
    e2e_ms = max(e0.elapsed_time(e1), e2e_wall_ms)  # host-synchronous API: wall clock is the honest number
    clocks = sampler.stop() if rank == 0 else None

    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([dev_ms, e2e_ms, float(sum(scan_ms) / max(len(scan_ms), 1))], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_ms, scan_avg = [float(x) for x in t.tolist()]
    else:
        scan_avg = sum(scan_ms) / max(len(scan_ms), 1)

    if rank == 0:
        ms_per_step = dev_ms / args.steps
        value = nq / (ms_per_step * 1e-3)
        e2e_value = nq / (e2e_ms / args.steps * 1e-3)
        peak, peak_src = _peaks()
        alg_bytes = n_local * dim * es + n_local * 4          # store rows once + cached inverse norms
        achieved = alg_bytes / (scan_avg * 1e-3) / 1e9 if scan_avg > 0 else 0.0
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak" if world == 1 else "strong", "vs_baseline": None,
            "dtype": "tf32" if (dt == "f32" and stats.scan_kernel == 2) else ("f32" if dt == "f32" else "bf16"),
            "data": "synthetic",
            "config": {"workload": f"{cfg}: {rows_total}x{dim} {dt} store, {nq}-query batch, top-{k} cosine, "
                                   f"{'row-sharded over %d GPUs + NCCL all-gather merge' % world if world > 1 else '1 GPU'}",
                       "rows_per_gpu": n_local,
                       "l2_policy": ("inputs larger than L2 (%.0f MB store per GPU vs 126 MB L2)" % (n_local * dim * es / 1e6))
                       if n_local * dim * es > 126e6 else
                       ("store (%.1f MB) fits L2: L2 flushed (256 MB write) before every timed step, steps timed one by one" % (n_local * dim * es / 1e6)),
                       "scan_kernel": {0: "exact_fp64", 1: "simt", 2: "tcgen05"}[int(stats.scan_kernel)],
                       "scan_ctas": int(stats.scan_ctas), "candidates_per_query": int(stats.candidates),
                       "exact_rescoring": "binary64, reference summation order (Neumaier)", "timing": "cuda events, max over ranks",
                       "exchange": ("peer memory (symmetric buffers, NVLink pull-merge kernel)" if peer_exchange else
                                    "ncclAllGather + merge kernel") if world > 1 else "none"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(Q.nbytes),
                    "d2h_bytes_per_step": int(nq * k * 16 + nq * 4 + 4), "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": _ncu_traffic(cfg, int(stats.scan_kernel), n_local), "kernel": "scan", "kernel_ms": scan_avg, "kernel_ms_samples": scan_calls, "algorithmic_bytes": alg_bytes,
                         "peak_source": peak_src},
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            qps, ms, desc = cpu_reference_leg(cfg, 1, 1, args.cpu_sample_rows or 100000)
            line["cpu_baseline"] = {"value": qps, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc}
            try:
                line["cpu_baseline"]["sklearn_blas_value"] = cpu_blas_leg(cfg, args.cpu_sample_rows or 100000)
            except Exception as e:  # pragma: no cover
                line["cpu_baseline"]["sklearn_blas_value"] = None
    store.close()
    if rank == 0:
        if world == 1 and cfg == "c2" and not args.rows and not args.no_scaling_baseline:
            # The multi-GPU runs use config C3 (100M x 384 bf16, strong scaling).  Its single-GPU point is
            # measured here so that the N = 2/4/8 lines have a same-workload N = 1 reference.
            try:
                r3, d3, dt3, nq3, k3, ss3, qs3 = CONFIGS["c3"]
                s3 = EmbeddingStore(d3, r3, dt3, device=local_rank)
                s3.synth_fill(ss3, r3)
                s3.set_size(r3)
                q3 = make_queries(s3, r3, nq3, d3, qs3, dev).to(dev)
                o3 = (torch.empty((nq3, k3), dtype=torch.int64, device=dev), torch.empty((nq3, k3), dtype=torch.float64, device=dev),
                      torch.empty((nq3,), dtype=torch.int32, device=dev))
                for _ in range(3):
                    s3.topk_device(q3, k3, out=o3, flags=vm.VM_FLAG_ASYNC)
                torch.cuda.synchronize()
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record()
                for _ in range(10):
                    s3.topk_device(q3, k3, out=o3, flags=vm.VM_FLAG_ASYNC)
                a1.record()
                torch.cuda.synchronize()
                ms3 = a0.elapsed_time(a1) / 10
                line["scaling_baseline"] = {"workload": f"c3: {r3}x{d3} {dt3} store on ONE GPU, {nq3}-query batch, top-{k3}",
                                            "value": nq3 / (ms3 * 1e-3), "unit": UNIT, "ms_per_step": ms3,
                                            "hbm_frac": (r3 * d3 * 2 + r3 * 4) / (ms3 * 1e-3) / 1e9 / _peaks()[0]}
                s3.close()
            except Exception as e:  # pragma: no cover - e.g. not enough free HBM
                line["scaling_baseline"] = {"error": repr(e)}
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


def _tensor_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json bf16_tflops_sustained: kernel timed inside a long step)"
        except Exception:
            pass
    return 1400.0, "fallback (B200_PROFILING.md sustained)"


def run_c4(args):
    """Config C4: all-pairs dedup, 1M x 768 bf16 vs itself, threshold 0.9 (planted near-duplicates)."""
    import numpy as np
    import torch
    import vidmem_b200 as vm
    from vidmem_b200 import dedup
    from vidmem_b200.store import EmbeddingStore
    rows, dim, dt, thr, seed, dup = C4
    if args.rows:
        rows = args.rows
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    st = EmbeddingStore(dim, rows, dt, device=local_rank)     # operand replicated on every rank (1.5 GB)
    st.synth_fill(seed, rows, dup_period=dup)
    torch.cuda.synchronize()
    x = st.rows[:rows]
    cap = 1 << 22
    lib = vm._lib.load()
    import ctypes as C
    oi = torch.empty((cap,), dtype=torch.int64, device=dev); oj = torch.empty_like(oi)
    os_ = torch.empty((cap,), dtype=torch.float32, device=dev); cnt = torch.zeros((1,), dtype=torch.int64, device=dev)

    def step():
        vm._lib.check(lib.vm_pairs_above(local_rank, x.data_ptr(), vm.VM_BF16, rows, dim, C.c_float(thr), cap, oi.data_ptr(),
                                         oj.data_ptr(), os_.data_ptr(), cnt.data_ptr(), rank, world, 0,
                                         torch.cuda.current_stream(dev).cuda_stream))

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    steps = max(1, min(args.steps, 5))
    for _ in range(max(1, min(args.warmup, 3))):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / steps
    hits_local = int(cnt.item())
    # e2e: rows come from pinned host memory every step, pairs go back to the host
    xh = torch.empty((rows, dim), dtype=torch.bfloat16).pin_memory()
    xh.copy_(x[:, :dim].cpu())
    xd = torch.empty_like(x)
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        xd[:, :dim].copy_(xh, non_blocking=True)
        i, j, s = dedup.pairs_above(xd, thr, cap=cap, part=rank, nparts=world)
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / steps
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms, e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms = [float(v) for v in t.tolist()]
        h = torch.tensor([hits_local], device=dev, dtype=torch.int64)
        dist.all_reduce(h)
        hits = int(h.item())
    else:
        hits = hits_local
    if rank == 0:
        pairs = rows * (rows - 1) / 2
        peak, src = _tensor_peak()
        tf = pairs * 2 * dim / (ms * 1e-3) / 1e12 / world      # per GPU
        line = {"metric": "dedup_unique_pairs_per_sec", "value": pairs / (ms * 1e-3), "unit": "pairs/s", "n_gpus": world,
                "steps": steps, "warmup": max(1, min(args.warmup, 3)), "ms_per_step": ms, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": f"c4: all-pairs dedup {rows}x{dim} bf16 vs itself, threshold {thr} (strict), "
                                       f"1/{dup} rows planted near-duplicates, upper triangle split over {world} GPU(s)",
                           "hits": hits, "l2_policy": "operand %.0f MB > 126 MB L2" % (rows * dim * 2 / 1e6)},
                "e2e": {"value": pairs / (e2e_ms * 1e-3), "unit": "pairs/s", "h2d_bytes_per_step": rows * dim * 2,
                        "d2h_bytes_per_step": hits_local * 20 + 8, "ms_per_step": e2e_ms},
                "gpu_launches": 3 * steps,
                "roofline": {"bound": "tensor", "achieved": tf, "peak": peak, "unit": "TFLOP/s", "frac": tf / peak,
                             "traffic": None, "kernel": "pairs_tc_kernel", "peak_source": src,
                             "flops_counted": "2*D per unordered pair (upper triangle only)"},
                "clocks": clocks}
        if world == 1 and not args.no_cpu_baseline:
            from oracle import oracle, synth
            ns = 4096
            E = synth.synth_rows(seed, 0, ns, dim, dup_period=dup)
            t0 = time.perf_counter()
            oracle.pairs_above(E, thr)
            dt_s = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": (ns * (ns - 1) / 2) / dt_s, "unit": "pairs/s", "cores": os.cpu_count() or 1,
                                    "kind": "port", "sample": f"{ns} x {dim} rows all-pairs (pairs/s is size-independent for the O(N^2 D) loop)"}
        print(json.dumps(line), flush=True)
    st.close()
    if world > 1:
        import torch.distributed as dist
        dist.barrier(); dist.destroy_process_group()


def run_c5(args):
    """Config C5: streaming mode -- 4096-row inserts interleaved with single-query top-10 lookups over a
    10M x 384 store on one GPU; reports p50/p99 latencies (host wall clock around synchronous calls)."""
    import numpy as np
    import torch
    import vidmem_b200 as vm
    from vidmem_b200.store import EmbeddingStore
    rows_total, dim, dt, nq, k, sseed, qseed = CONFIGS["c5"]
    if args.rows:
        rows_total = args.rows
    dt = args.c5_dtype or dt
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    rounds = max(8, min(args.steps, 64))
    per_round_q = 8
    ins = 4096
    st = EmbeddingStore(dim, rows_total + rounds * ins, dt, device=0)
    st.synth_fill(sseed, rows_total)
    st.set_size(rows_total)
    torch.cuda.synchronize()
    qpin = make_queries(st, rows_total, rounds * per_round_q, dim, qseed, dev).pin_memory().numpy()
    g = torch.Generator(device="cpu").manual_seed(sseed + 1)
    new_rows = (torch.randint(-127, 128, (rounds * ins, dim), generator=g).float() / 128.0).pin_memory()
    for w in range(3):
        st.topk(qpin[w:w + 1], k)
    sampler = ClockSampler(0)
    sampler.start()
    q_lat, i_lat = [], []
    t_all = time.perf_counter()
    for r in range(rounds):
        t0 = time.perf_counter()
        st.append(new_rows[r * ins:(r + 1) * ins])          # host -> device, convert, norms; synchronous
        i_lat.append((time.perf_counter() - t0) * 1e3)
        for qi in range(per_round_q):
            t0 = time.perf_counter()
            st.topk(qpin[r * per_round_q + qi:r * per_round_q + qi + 1], k)
            q_lat.append((time.perf_counter() - t0) * 1e3)
    wall = time.perf_counter() - t_all
    clocks = sampler.stop()
    scan_ms = []
    for qi in range(5):
        st.topk(qpin[qi:qi + 1], k, flags=vm.VM_FLAG_TIMING)
        scan_ms.append(st.last_scan_ms())
    es = 4 if dt == "f32" else 2
    n_now = len(st)
    peak, src = _peaks()
    alg = n_now * dim * es + n_now * 4
    scan = sum(scan_ms) / len(scan_ms)
    pct = lambda a, p: float(np.percentile(np.asarray(a), p))
    line = {"metric": "streaming_queries_per_sec_top10_384d", "value": len(q_lat) / (sum(q_lat) * 1e-3), "unit": "queries/s",
            "n_gpus": 1, "steps": rounds, "warmup": 3, "ms_per_step": wall * 1e3 / rounds, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "tf32" if dt == "f32" else "bf16", "data": "synthetic",
            "config": {"workload": f"c5: streaming, {rows_total}x{dim} {dt} store, {ins}-row inserts interleaved with "
                                   f"{per_round_q} single-query top-{k} lookups per insert, host buffers, synchronous calls",
                       "scan_kernel": {0: "exact_fp64", 1: "simt", 2: "tcgen05"}[int(st.last_stats.scan_kernel)]},
            "latency_ms": {"query_p50": pct(q_lat, 50), "query_p99": pct(q_lat, 99), "insert_p50": pct(i_lat, 50),
                           "insert_p99": pct(i_lat, 99)},
            "e2e": {"value": len(q_lat) / (sum(q_lat) * 1e-3), "unit": "queries/s", "h2d_bytes_per_step": ins * dim * 4 + per_round_q * dim * 4,
                    "d2h_bytes_per_step": per_round_q * (k * 16 + 8)},
            "gpu_launches": int(st.last_stats.scan_launches) * len(q_lat) + 2 * rounds,
            "roofline": {"bound": "hbm", "achieved": alg / (scan * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": alg / (scan * 1e-3) / 1e9 / peak, "traffic": None, "kernel": "scan", "kernel_ms": scan,
                         "algorithmic_bytes": alg, "peak_source": src},
            "clocks": clocks}
    print(json.dumps(line), flush=True)
    st.close()


def torchrun_argv(gpus: int, argv, port: int):
    """The launch line of the bench contract for N > 1 (one rank per GPU, 127.0.0.1 rendezvous)."""
    return [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={gpus}", "--master-addr", "127.0.0.1",
            "--master-port", str(port), os.path.abspath(__file__), *argv]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default=None, choices=sorted(CONFIGS) + ["c4"])
    ap.add_argument("--rows", type=int, default=None, help="override the total row count (debug)")
    ap.add_argument("--flags", type=int, default=0, help="extra VM_FLAG_* bits (debug: 4 = force SIMT, 8 = force tcgen05)")
    ap.add_argument("--cpu-sample-rows", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--peer-exchange", action="store_true",
                    help="multi-GPU: replace ncclAllGather + merge by the peer-memory pull-merge kernel (symmetric memory)")
    ap.add_argument("--no-scaling-baseline", action="store_true", help="skip the extra C3-on-one-GPU measurement of the N=1 run")
    ap.add_argument("--c5-dtype", default=None, choices=["f32", "bf16"])
    args = ap.parse_args()
    if args.gpus > 1 and args.impl == "ours" and "WORLD_SIZE" not in os.environ:
        # started as plain `python bench.py --gpus N`: re-launch under torchrun, one rank per GPU
        import socket
        with socket.socket() as sk:
            sk.bind(("127.0.0.1", 0))
            port = sk.getsockname()[1]
        os.execv(sys.executable, torchrun_argv(args.gpus, sys.argv[1:], port))
    if args.impl == "reference":
        run_reference(args)
    elif args.config == "c4":
        run_c4(args)
    elif args.config == "c5":
        run_c5(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
