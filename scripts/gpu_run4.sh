#!/bin/bash
# round 2, call 4: select_rescore slab loading (warp per slab, batched loads): tests, C2 iid/clustered, ncu of the band-settling select
O=gpurun_out/r2e; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.txt
tail -5 $O/pytest_gpu.txt
timeout 300 python bench.py --steps 20 --warmup 5 --only-main --no-scaling-baseline --no-cpu-baseline > $O/bench_c2.json 2> $O/bench_c2.err; echo "c2 rc=$?"
for rho in 0.05 0.2; do
timeout 300 python bench.py --steps 20 --warmup 5 --only-main --no-scaling-baseline --variant clustered --cluster-rho $rho --no-cpu-baseline > $O/bench_c2_clustered_rho$rho.json 2> $O/bench_c2_clustered_rho$rho.err; echo "c2 clustered rho=$rho rc=$?"
done
timeout 300 python bench.py --steps 20 --warmup 5 --only-main --config c3 --rows 12500000 --no-cpu-baseline --variant clustered > $O/bench_c3shard_clustered.json 2> $O/bench_c3shard_clustered.err; echo "c3 shard clustered rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_c2_rho0.05.csv \
    python bench.py --steps 2 --warmup 3 --only-main --no-cpu-baseline --no-scaling-baseline --variant clustered --cluster-rho 0.05 > $O/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
ncu --set full --import-source on --clock-control none -k regex:select_rescore_kernel -s 6 -c 1 -f -o $O/select_c2_rho0.05 python bench.py --steps 4 --warmup 3 --only-main --no-cpu-baseline --no-scaling-baseline --variant clustered --cluster-rho 0.05 > $O/ncu_select.log 2>&1; echo "ncu select rc=$?"
tail -c 300 $O/*.err
