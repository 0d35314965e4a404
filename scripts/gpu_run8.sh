#!/bin/bash
# round 2, call 8: full GPU suite; A/B of programmatic dependent launch inside the captured graph (host-buffer path)
O=gpurun_out/r2i; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.txt
tail -6 $O/pytest_gpu.txt
for rep in 1 2 3; do
for lib in real-time-brain-inspired-video-memory_b200/libvidmem.so _ab/libvidmem_gpdl.so; do
  tag=$(basename $lib .so)
  VIDMEM_LIB=$PWD/$lib timeout 300 python bench.py --steps 30 --warmup 5 --only-main --no-scaling-baseline --no-cpu-baseline > $O/ab_c2_${tag}_$rep.json 2> $O/ab_c2_${tag}_$rep.err
  VIDMEM_LIB=$PWD/$lib timeout 300 python bench.py --steps 30 --warmup 5 --only-main --config c1 --no-cpu-baseline > $O/ab_c1_${tag}_$rep.json 2> $O/ab_c1_${tag}_$rep.err
done; done
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2i/ab_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f.split('/')[-1], round(d['value']), round(d['e2e']['value']), round(d['e2e']['ms_per_step']*1e3,1),'us e2e')
    except Exception as e: print(f,'ERR',e)
P
tail -c 300 $O/*.err
