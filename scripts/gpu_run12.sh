#!/bin/bash
# round 2, evidence run (1 GPU): complete default bench line, reference arm, launch lists, ncu --set full of the dominant kernels
O=gpurun_out/r2m; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt
timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
tail -c 600 $O/bench_n1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"
timeout 300 python bench.py --steps 30 --warmup 5 --only-main --config c1 > $O/bench_c1.json 2> $O/bench_c1.err; echo "c1 rc=$?"
timeout 300 python bench.py --steps 20 --warmup 5 --only-main --config c3 --rows 12500000 --no-cpu-baseline > $O/bench_c3shard.json 2> $O/bench_c3shard.err; echo "c3 shard rc=$?"
timeout 300 python bench.py --steps 20 --warmup 5 --only-main --no-scaling-baseline --no-cpu-baseline --variant clustered --cluster-rho 0.05 > $O/bench_c2_tight.json 2> $O/bench_c2_tight.err; echo "tight rc=$?"
# launch lists (device time per launch, cold cache, serialised): main C2 step, C4 at 131 072 rows
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'scan_tc|select_rescore|normalize_queries|exact_scan|pairs' -c 60 --csv --log-file $O/launches_c2.csv \
    python bench.py --steps 2 --warmup 3 --only-main --no-cpu-baseline --no-scaling-baseline > $O/ncu_launch_c2.log 2>&1; echo "ncu launches c2 rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'pairs' -c 20 --csv --log-file $O/launches_c4_131k.csv \
    python bench.py --config c4 --rows 131072 --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_launch_c4.log 2>&1; echo "ncu launches c4 rc=$?"
# full captures
ncu --set full --import-source on --clock-control none -k regex:scan_tc_kernel -s 6 -c 1 -f -o $O/scan_c2_f32 python bench.py --steps 4 --warmup 3 --only-main --no-cpu-baseline --no-scaling-baseline > $O/ncu_scan_c2.log 2>&1; echo "ncu scan c2 rc=$?"
ncu --set full --import-source on --clock-control none -k regex:select_rescore_kernel -s 6 -c 1 -f -o $O/select_c2_f32 python bench.py --steps 4 --warmup 3 --only-main --no-cpu-baseline --no-scaling-baseline > $O/ncu_select_c2.log 2>&1; echo "ncu select c2 rc=$?"
ncu --set full --clock-control none -k regex:scan_tc_kernel -s 2 -c 1 -f -o $O/scan_bf16_12500000 python scripts/ncu_scan_shape.py 12500000 bf16 > $O/ncu_scan_bf16.log 2>&1; echo "ncu scan bf16 rc=$?"
ncu --set full --clock-control none -k regex:pairs_tc2 -c 1 -f -o $O/pairs_c4_131k python bench.py --config c4 --rows 131072 --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_pairs.log 2>&1; echo "ncu pairs rc=$?"
ls -la $O
