#!/bin/bash
O=gpurun_out/r2u; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.txt
tail -6 $O/pytest_gpu.txt
timeout 300 python bench.py --steps 20 --warmup 5 --only-main --no-scaling-baseline --no-cpu-baseline > $O/bench_c2.json 2> $O/bench_c2.err; echo "c2 rc=$?"
timeout 300 python bench.py --steps 20 --warmup 5 --only-main --config c3 --rows 12500000 --no-cpu-baseline > $O/bench_c3shard.json 2> $O/bench_c3shard.err; echo "c3 shard rc=$?"
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2u/bench_*.json')):
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f.split('/')[-1], round(d['value']), d['ms_per_step'], d['parity'])
P
