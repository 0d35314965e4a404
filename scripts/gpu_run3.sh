#!/bin/bash
O=gpurun_out/r2d; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.txt
tail -15 $O/pytest_gpu.txt
timeout 300 python bench.py --steps 20 --warmup 5 --only-main --no-scaling-baseline > $O/bench_c2.json 2> $O/bench_c2.err; echo "c2 rc=$?"
timeout 300 python bench.py --steps 20 --warmup 5 --only-main --no-scaling-baseline --variant clustered --no-cpu-baseline > $O/bench_c2_clustered.json 2> $O/bench_c2_clustered.err; echo "c2 clustered rc=$?"
for rho in 0.05 0.5; do
timeout 300 python bench.py --steps 20 --warmup 5 --only-main --no-scaling-baseline --variant clustered --cluster-rho $rho --no-cpu-baseline > $O/bench_c2_clustered_rho$rho.json 2> $O/bench_c2_clustered_rho$rho.err; echo "c2 clustered rho=$rho rc=$?"
done
timeout 300 python bench.py --steps 20 --warmup 5 --only-main --config c3 --rows 12500000 --no-cpu-baseline > $O/bench_c3shard.json 2> $O/bench_c3shard.err; echo "c3 shard rc=$?"
timeout 300 python bench.py --steps 20 --warmup 5 --only-main --config c3 --rows 12500000 --no-cpu-baseline --variant clustered > $O/bench_c3shard_clustered.json 2> $O/bench_c3shard_clustered.err; echo "c3 shard clustered rc=$?"
ncu --set full --import-source on --clock-control none -k regex:select_rescore_kernel -s 6 -c 1 -f -o $O/select_c2 python bench.py --steps 4 --warmup 3 --only-main --no-cpu-baseline --no-scaling-baseline > $O/ncu_select.log 2>&1; echo "ncu select rc=$?"
tail -c 300 $O/*.err
