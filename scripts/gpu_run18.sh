#!/bin/bash
O=gpurun_out/r2s; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.txt
tail -6 $O/pytest_gpu.txt
timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
tail -c 600 $O/bench_n1.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2s/bench_n1.json').read().strip().splitlines()[-1])
print('main', round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), d['parity']['ok'])
for k in ('clustered','dedup','streaming','streaming_bf16'):
    r=d[k]; print(k, r['value'], r['ms_per_step'], (r.get('parity') or {}).get('ok'), r.get('latency_ms'))
for k,r in d['binary64_store'].items(): print(k, r['value'], r['parity']['ok'])
print(d['dedup']['roofline'])
P
