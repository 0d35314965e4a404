#!/bin/bash
# One-GPU evidence run -- what profiles/ is made from.  Run on the GPU box from the repo root:
#     gpurun --timeout 1500 -- 'scripts/gpu_evidence.sh gpurun_out/<tag> [quick]'
# quick = tests + smoke + both bench arms only (no ncu).  Numbers printed under ncu are never bench values: every ncu pass
# below re-runs a command that has already exited 0 without it.
O=${1:-gpurun_out/evidence}; MODE=${2:-full}; mkdir -p $O
B="python bench.py --only-main --no-cpu-baseline --no-scaling-baseline"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.txt
tail -4 $O/pytest_gpu.txt
timeout 300 python __graft_entry__.py smoke > $O/smoke.txt 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "reference arm rc=$?"
[ "$MODE" = quick ] && exit 0
# sub-configurations on their own
timeout 300 python bench.py --steps 20 --warmup 5 --only-main --config c3 --rows 12500000 --no-cpu-baseline > $O/bench_c3shard.json 2> $O/bench_c3shard.err; echo "c3 shard rc=$?"
timeout 300 python bench.py --config c4 --steps 3 --warmup 1 > $O/bench_c4.json 2> $O/bench_c4.err; echo "c4 rc=$?"
# launch lists (device time per launch: cold cache, serialised -- shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'scan_tc|select_rescore|normalize_queries|exact_scan' -c 60 --csv \
    --log-file $O/launches_c2.csv $B --steps 2 --warmup 3 > $O/ncu_launch_c2.log 2>&1; echo "launch list c2 rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'pairs' -c 20 --csv --log-file $O/launches_c4_262k.csv \
    python bench.py --config c4 --rows 262144 --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_launch_c4.log 2>&1; echo "launch list c4 rc=$?"
# full captures of the dominant kernels (one launch each; read here with scripts/ncu_summary.py)
ncu --set full --import-source on --clock-control none -k regex:scan_tc_kernel -s 6 -c 1 -f -o $O/scan_c2_f32 $B --steps 4 --warmup 3 > $O/ncu_scan_c2.log 2>&1; echo "ncu scan c2 rc=$?"
ncu --set full --import-source on --clock-control none -k regex:select_rescore_kernel -s 6 -c 1 -f -o $O/select_c2_f32 $B --steps 4 --warmup 3 > $O/ncu_select_c2.log 2>&1; echo "ncu select c2 rc=$?"
ncu --set full --clock-control none -k regex:scan_tc_kernel -s 2 -c 1 -f -o $O/scan_bf16_12500000 python scripts/ncu_scan_shape.py 12500000 bf16 > $O/ncu_scan_bf16.log 2>&1; echo "ncu scan bf16 rc=$?"
ncu --set full --clock-control none -k regex:pairs_tc2 -c 1 -f -o $O/pairs_c4_262k python bench.py --config c4 --rows 262144 --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_pairs.log 2>&1; echo "ncu pairs rc=$?"
ls -la $O
