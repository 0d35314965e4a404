#!/bin/bash
# round 2, 2 GPUs: the collective tests + the full N=2 bench line (C3 sharded + clustered + dedup + streaming, in-run parity)
O=gpurun_out/r2k; mkdir -p $O
nvidia-smi -L > $O/smi.txt
timeout 600 python -m pytest tests/test_sharded_gpu.py -x -q > $O/pytest_sharded.txt 2>&1; echo "pytest sharded rc=$?" >> $O/pytest_sharded.txt
tail -5 $O/pytest_sharded.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench n2 rc=$?"
tail -c 1500 $O/bench_n2.err
python - <<'P'
import json
try:
    d=json.loads(open('gpurun_out/r2k/bench_n2.json').read().strip().splitlines()[-1])
    print('c3 n2', round(d['value']), d['ms_per_step'], d['parity'], 'e2e', d['e2e']['value'])
    for k in ("clustered","clustered_tight","dedup","streaming","streaming_bf16"):
        r=d.get(k) or {}
        print(k, r.get('value'), r.get('ms_per_step'), r.get('parity'))
    print('cpu_baseline', d.get('cpu_baseline'))
except Exception as e: print('ERR', e)
P
