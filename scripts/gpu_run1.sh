#!/bin/bash
# round 2, GPU call 1: tests + complete bench line + launch list + traffic captures for the N=2/4 shard sizes
mkdir -p gpurun_out/r2a
O=gpurun_out/r2a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.txt
tail -5 $O/pytest_gpu.txt
python bench.py --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
tail -c 600 $O/bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_c2.csv \
    python bench.py --steps 2 --warmup 1 --only-main --no-cpu-baseline --no-scaling-baseline > $O/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
for shape in "25000000 bf16" "50000000 bf16"; do
  set -- $shape
  ncu --set full --clock-control none -k regex:scan_tc_kernel -s 2 -c 1 -f -o $O/scan_$2_$1 python scripts/ncu_scan_shape.py $1 $2 > $O/ncu_scan_$2_$1.log 2>&1; echo "ncu $shape rc=$?"
done
ls -la $O
