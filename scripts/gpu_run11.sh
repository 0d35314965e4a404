#!/bin/bash
# round 2, call 11: own-k-th service variants on a tight-cluster store (band keys kept per query, scan time)
O=gpurun_out/r2l; mkdir -p $O
for lib in real-time-brain-inspired-video-memory_b200/libvidmem.so _ab/libvidmem_l4r0.so _ab/libvidmem_l4r1.so _ab/libvidmem_l2r0.so; do
  tag=$(basename $lib .so)
  for v in "iid" "clustered --cluster-rho 0.2" "clustered --cluster-rho 0.05" "clustered --cluster-rho 0.02"; do
    name=$(echo $v | tr ' ' '_' | tr -d '-')
    VIDMEM_LIB=$PWD/$lib timeout 300 python bench.py --steps 20 --warmup 5 --only-main --no-scaling-baseline --no-cpu-baseline --variant $v > $O/${tag}_$name.json 2> $O/${tag}_$name.err
  done
  VIDMEM_LIB=$PWD/$lib timeout 300 python bench.py --steps 20 --warmup 5 --only-main --config c3 --rows 12500000 --no-cpu-baseline --variant clustered --cluster-rho 0.05 > $O/${tag}_c3shard_rho0.05.json 2> $O/${tag}_c3shard_rho0.05.err
done
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2l/*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']; c=d['certification']
        print(f.split('/')[-1], round(d['value']), round(d['ms_per_step'],4), 'scan', round(r['kernel_ms'],4), 'kept', c.get('band_keys_per_query'), 'spilled', c.get('spilled_keys_per_query'), 'unc', c['uncertified'], 'band', c['band_settled'], d['parity']['ok'])
    except Exception as e: print(f,'ERR',e)
P
tail -c 300 $O/*.err
