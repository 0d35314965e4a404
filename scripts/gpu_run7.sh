#!/bin/bash
# round 2, call 7: band batches staged through shared memory; binary64 bench variants; smoke
O=gpurun_out/r2h; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.txt
tail -6 $O/pytest_gpu.txt
timeout 300 python __graft_entry__.py smoke > $O/smoke.txt 2>&1; echo "smoke rc=$?"; tail -4 $O/smoke.txt
timeout 300 python bench.py --steps 20 --warmup 5 --only-main --no-scaling-baseline --no-cpu-baseline > $O/bench_c2.json 2> $O/bench_c2.err; echo "c2 rc=$?"
for rho in 0.05 0.2; do
timeout 300 python bench.py --steps 20 --warmup 5 --only-main --no-scaling-baseline --variant clustered --cluster-rho $rho --no-cpu-baseline > $O/bench_c2_clustered_rho$rho.json 2> $O/bench_c2_clustered_rho$rho.err; echo "c2 clustered rho=$rho rc=$?"
done
timeout 300 python bench.py --steps 20 --warmup 5 --only-main --config c3 --rows 12500000 --no-cpu-baseline --variant clustered > $O/bench_c3shard_clustered.json 2> $O/bench_c3shard_clustered.err; echo "c3 shard clustered rc=$?"
for sd in f64 f64+bf16 bf16; do
timeout 300 python bench.py --steps 20 --warmup 5 --only-main --no-scaling-baseline --store-dtype $sd > $O/bench_c2_$sd.json 2> $O/bench_c2_$sd.err; echo "c2 $sd rc=$?"
done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'scan_tc|select_rescore|normalize_queries|exact_scan' -c 40 --csv --log-file $O/launches_c2_rho0.05.csv \
    python bench.py --steps 2 --warmup 3 --only-main --no-cpu-baseline --no-scaling-baseline --variant clustered --cluster-rho 0.05 > $O/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
tail -c 600 $O/*.err
