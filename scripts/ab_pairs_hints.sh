#!/bin/bash
# Same-box A/B of all-pairs kernel builds at the full C4 size (libraries under _ab/ built with
# build.py -D... --out=_ab/libvidmem_<tag>.so; the shipped library is the first arm):
#     gpurun --timeout 300 -- 'scripts/ab_pairs_hints.sh gpurun_out/<tag> v2 v2p'
O=${1:-gpurun_out/ab_pairs}; shift; mkdir -p $O
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.avg.per_second,lts__t_sector_hit_rate.pct
CUR=$PWD/real-time-brain-inspired-video-memory_b200/libvidmem.so
LIBS="$CUR"; for v in "$@"; do LIBS="$LIBS $PWD/_ab/libvidmem_$v.so"; done
for lib in $LIBS; do
  VIDMEM_LIB=$lib timeout 120 python scripts/ab_pairs.py 1000000 3 2>&1 | tail -1 | tee -a $O/timing.txt
done
for lib in $LIBS; do
  t=$(basename $lib .so)
  VIDMEM_LIB=$lib timeout 120 ncu --metrics $M --clock-control none -k regex:pairs_tc2 -s 1 -c 1 --csv --log-file $O/pairs_1m_$t.csv \
      python scripts/ab_pairs.py 1000000 2 > $O/ncu_$t.log 2>&1; echo "ncu $t rc=$?"
  grep '^"' $O/pairs_1m_$t.csv | awk -F'","' '{print $(NF-2), $(NF)}' | tr -d '"' | tr '\n' ';'; echo
done
