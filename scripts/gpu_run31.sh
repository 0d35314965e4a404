#!/bin/bash
# round 2, 8 GPUs: the complete N=8 bench line (C3 + clustered + dedup + streaming, in-run parity), then the reference arm
O=gpurun_out/r2w; mkdir -p $O
nvidia-smi -L > $O/smi.txt
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --steps 20 --warmup 5 > $O/bench_n4.json 2> $O/bench_n4.err; echo "bench n8 rc=$?"
tail -c 800 $O/bench_n4.err
python - <<'P'
import json
try:
    d=json.loads(open('gpurun_out/r2w/bench_n4.json').read().strip().splitlines()[-1])
    print('c3 n8', round(d['value']), d['ms_per_step'], 'scan', d['roofline']['kernel_ms'], d['roofline']['frac'], d['parity']['ok'], 'e2e', d['e2e']['value'])
    for k in ('clustered','dedup','streaming'):
        r=d.get(k) or {}
        print(k, r.get('value'), r.get('ms_per_step'), (r.get('parity') or {}).get('ok'))
except Exception as e: print('ERR', e)
P
