#!/bin/bash
O=gpurun_out/r2c; mkdir -p $O
CUR=real-time-brain-inspired-video-memory_b200/libvidmem.so
for shape in "10000000 bf16 1" "12500000 bf16 64"; do
  for lib in _ab/libvidmem_r1.so $CUR _ab/libvidmem_NOSPLITCODE.so _ab/libvidmem_NOEPI.so; do
    timeout 120 python scripts/ab_scan.py $lib $shape 32 20 2>&1 | tail -1
  done
done | tee $O/ab3.txt
