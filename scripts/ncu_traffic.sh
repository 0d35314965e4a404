#!/bin/bash
# Per-launch DRAM traffic / duration / tensor-pipe activity of the dominant kernel at the shapes whose bench records would
# otherwise carry "traffic": null -- one metrics-only ncu pass per shape (a few replays, not --set full):
#     gpurun --timeout 400 -- 'scripts/ncu_traffic.sh gpurun_out/<tag>'
# scripts/ncu_metrics_summary.py turns each CSV into a profiles/*_summary.json that bench.py's roofline.traffic reads.
O=${1:-gpurun_out/traffic}; mkdir -p $O
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.avg.per_second,lts__t_sector_hit_rate.pct
scan() {  # rows dtype nq tag
  timeout 240 ncu --metrics $M --clock-control none -k regex:scan_tc_kernel -s 2 -c 1 --csv --log-file $O/scan_$4.csv \
      python scripts/ncu_scan_shape.py $1 $2 384 $3 > $O/scan_$4.log 2>&1; echo "scan $4 rc=$?"
}
timeout 300 ncu --metrics $M --clock-control none -k regex:pairs_tc2 -s 1 -c 1 --csv --log-file $O/pairs_c4_1m.csv \
    python bench.py --config c4 --steps 1 --warmup 1 --no-cpu-baseline > $O/pairs_c4_1m.log 2>&1; echo "pairs 1M rc=$?"
scan 10000000 f32 1 c5_f32
scan 10000000 bf16 1 c5_bf16
scan 100000000 bf16 64 c3_n1
scan 5000 f32 30 c1
ls -la $O
