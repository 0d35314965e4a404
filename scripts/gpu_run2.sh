#!/bin/bash
# round 2, GPU call 2: band-complete scan + PDL chain -- tests first, then short benches and A/B of the bf16 split
O=gpurun_out/r2b; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.txt
tail -25 $O/pytest_gpu.txt
timeout 300 python bench.py --steps 20 --warmup 5 --only-main --no-scaling-baseline > $O/bench_c2.json 2> $O/bench_c2.err; echo "c2 rc=$?"
timeout 300 python bench.py --steps 20 --warmup 5 --only-main --no-scaling-baseline --variant clustered --no-cpu-baseline > $O/bench_c2_clustered.json 2> $O/bench_c2_clustered.err; echo "c2 clustered rc=$?"
timeout 300 python bench.py --steps 20 --warmup 5 --only-main --config c1 > $O/bench_c1.json 2> $O/bench_c1.err; echo "c1 rc=$?"
for fl in 0 32; do
timeout 300 python bench.py --steps 20 --warmup 5 --only-main --config c3 --rows 12500000 --no-cpu-baseline --flags $fl > $O/bench_c3shard_flags$fl.json 2> $O/bench_c3shard_flags$fl.err; echo "c3 shard flags=$fl rc=$?"
timeout 300 python bench.py --steps 20 --warmup 5 --only-main --config c3 --rows 12500000 --no-cpu-baseline --variant clustered --flags $fl > $O/bench_c3shard_clustered_flags$fl.json 2> $O/bench_c3shard_clustered_flags$fl.err; echo "c3 shard clustered flags=$fl rc=$?"
done
timeout 300 python bench.py --steps 20 --config c5 > $O/bench_c5.json 2> $O/bench_c5.err; echo "c5 rc=$?"
timeout 300 python bench.py --steps 20 --config c5 --c5-dtype bf16 > $O/bench_c5_bf16.json 2> $O/bench_c5_bf16.err; echo "c5 bf16 rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_c2.csv \
    python bench.py --steps 2 --warmup 1 --only-main --no-cpu-baseline --no-scaling-baseline > $O/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
tail -c 400 $O/*.err
for nq in 16 32 48; do for fl in 0 32; do timeout 120 python scripts/ab_scan.py real-time-brain-inspired-video-memory_b200/libvidmem.so 12500000 bf16 $nq $fl 20 2>&1 | tail -1; done; done | tee $O/ab_split_nq.txt
