#!/bin/bash
O=gpurun_out/r2z; mkdir -p $O
timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2z/bench_n1.json').read().strip().splitlines()[-1])
print('main', round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), d['parity']['ok'])
for k in ('clustered','clustered_tight'):
    r=d[k]; print(k, r['value'], r['ms_per_step'], r['parity']['ok'], r['certification'])
for k in ('dedup','streaming','streaming_bf16'):
    r=d[k]; print(k, r['value'], r['ms_per_step'], (r.get('parity') or {}).get('ok'), r.get('latency_ms'))
P
tail -c 300 $O/bench_n1.err
