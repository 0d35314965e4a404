#!/bin/bash
# round 2, call 5: binary64 stores, growable stores, band_topk rewrite
O=gpurun_out/r2f; mkdir -p $O
timeout 600 python -m pytest tests/test_exact_store_gpu.py tests/test_growable_gpu.py -x -q > $O/pytest_new.txt 2>&1; echo "pytest new rc=$?" >> $O/pytest_new.txt
tail -30 $O/pytest_new.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.txt
tail -8 $O/pytest_gpu.txt
timeout 300 python bench.py --steps 20 --warmup 5 --only-main --no-scaling-baseline --no-cpu-baseline > $O/bench_c2.json 2> $O/bench_c2.err; echo "c2 rc=$?"
for rho in 0.05; do
timeout 300 python bench.py --steps 20 --warmup 5 --only-main --no-scaling-baseline --variant clustered --cluster-rho $rho --no-cpu-baseline > $O/bench_c2_clustered_rho$rho.json 2> $O/bench_c2_clustered_rho$rho.err; echo "c2 clustered rho=$rho rc=$?"
done
timeout 300 python bench.py --steps 20 --warmup 5 --only-main --config c3 --rows 12500000 --no-cpu-baseline --variant clustered > $O/bench_c3shard_clustered.json 2> $O/bench_c3shard_clustered.err; echo "c3 shard clustered rc=$?"
timeout 300 python bench.py --steps 24 --config c5 > $O/bench_c5.json 2> $O/bench_c5.err; echo "c5 rc=$?"
tail -c 600 $O/*.err
