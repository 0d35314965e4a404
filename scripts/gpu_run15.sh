#!/bin/bash
# round 2, call 15: raster group width of the all-pairs kernel (8 / 16 / 32 / 64 column blocks per group); new binary64-store tests
O=gpurun_out/r2p; mkdir -p $O
timeout 900 python -m pytest tests/test_exact_store_gpu.py tests/test_adapters_gpu.py -x -q > $O/pytest_new.txt 2>&1; echo "pytest new rc=$?" >> $O/pytest_new.txt
tail -6 $O/pytest_new.txt
for lib in real-time-brain-inspired-video-memory_b200/libvidmem.so _ab/libvidmem_gj4.so _ab/libvidmem_gj5.so _ab/libvidmem_gj6.so; do
  tag=$(basename $lib .so)
  VIDMEM_LIB=$PWD/$lib timeout 300 python -m pytest tests/test_pairs_gpu.py -x -q 2>&1 | tail -1
  VIDMEM_LIB=$PWD/$lib timeout 300 python bench.py --config c4 --steps 3 --warmup 1 --no-cpu-baseline > $O/c4_$tag.json 2> $O/c4_$tag.err
  VIDMEM_LIB=$PWD/$lib timeout 300 python bench.py --config c4 --rows 131072 --steps 5 --warmup 2 --no-cpu-baseline > $O/c4_131k_$tag.json 2> $O/c4_131k_$tag.err
done
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2p/c4_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f.split('/')[-1], '%.4g'%d['value'], round(d['ms_per_step'],2), 'TF', round(d['roofline']['achieved'],1), d['clocks']['sm_mhz'], d['parity']['ok'], d['config'].get('hits'))
    except Exception as e: print(f,'ERR',e)
P
tail -c 300 $O/*.err
