#!/bin/bash
O=gpurun_out/r2x; mkdir -p $O
timeout 600 python -m pytest tests/test_growable_gpu.py -x -q 2>&1 | tail -2
for rep in 1 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 2955$rep bench.py --gpus 4 --config c5 --steps 20 > $O/c5_n4_$rep.json 2> $O/c5_n4_$rep.err; echo "c5 n4 rc=$?"
done
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2x/c5_n4_*.json')):
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['value'], d['latency_ms'], d['growth']['capacity_rows_before'], d['growth']['capacity_rows_after'])
P
tail -c 300 $O/*.err
