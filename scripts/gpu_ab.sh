#!/bin/bash
O=gpurun_out/r2c; mkdir -p $O
CUR=real-time-brain-inspired-video-memory_b200/libvidmem.so
for shape in "10000000 bf16 1" "12500000 bf16 64" "1000000 bf16 64" "1000000 f32 64" "10000000 f32 1"; do
  timeout 120 python scripts/ab_scan.py _ab/libvidmem_r1.so $shape 0 20 2>&1 | tail -1
  timeout 120 python scripts/ab_scan.py $CUR $shape 0 20 2>&1 | tail -1
  timeout 120 python scripts/ab_scan.py $CUR $shape 32 20 2>&1 | tail -1
done | tee $O/ab4.txt
