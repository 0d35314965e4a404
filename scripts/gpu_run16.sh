#!/bin/bash
O=gpurun_out/r2q; mkdir -p $O
for lib in _ab/libvidmem_gj6.so _ab/libvidmem_gj7.so _ab/libvidmem_gj8.so; do
  tag=$(basename $lib .so)
  VIDMEM_LIB=$PWD/$lib timeout 300 python -m pytest tests/test_pairs_gpu.py -x -q 2>&1 | tail -1
  VIDMEM_LIB=$PWD/$lib timeout 300 python bench.py --config c4 --steps 3 --warmup 1 --no-cpu-baseline > $O/c4_$tag.json 2> $O/c4_$tag.err
  VIDMEM_LIB=$PWD/$lib timeout 300 python bench.py --config c4 --rows 131072 --steps 10 --warmup 3 --no-cpu-baseline > $O/c4_131k_$tag.json 2> $O/c4_131k_$tag.err
  VIDMEM_LIB=$PWD/$lib timeout 300 python bench.py --config c4 --rows 32768 --steps 20 --warmup 3 --no-cpu-baseline > $O/c4_32k_$tag.json 2> $O/c4_32k_$tag.err
done
VIDMEM_LIB=$PWD/real-time-brain-inspired-video-memory_b200/libvidmem.so timeout 300 python bench.py --config c4 --rows 32768 --steps 20 --warmup 3 --no-cpu-baseline > $O/c4_32k_libvidmem.json 2> $O/c4_32k_libvidmem.err
VIDMEM_LIB=$PWD/real-time-brain-inspired-video-memory_b200/libvidmem.so timeout 300 python bench.py --config c4 --rows 131072 --steps 10 --warmup 3 --no-cpu-baseline > $O/c4_131k_libvidmem.json 2> $O/c4_131k_libvidmem.err
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2q/c4_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f.split('/')[-1], '%.4g'%d['value'], round(d['ms_per_step'],3), 'TF', round(d['roofline']['achieved'],1), d['clocks']['sm_mhz'], d['parity']['ok'], d['config'].get('hits'))
    except Exception as e: print(f,'ERR',e)
P
tail -c 300 $O/*.err
