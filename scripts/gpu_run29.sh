#!/bin/bash
O=gpurun_out/r2v; mkdir -p $O
timeout 600 python -m pytest tests/test_pairs_gpu.py tests/test_sharded_gpu.py -x -q > $O/pytest.txt 2>&1; echo "pytest rc=$?" >> $O/pytest.txt
tail -4 $O/pytest.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --config c4 --steps 3 --warmup 1 --no-cpu-baseline > $O/c4_n2.json 2> $O/c4_n2.err; echo "c4 n2 rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2v/c4_n2.json').read().strip().splitlines()[-1])
print('c4 n2', '%.4g'%d['value'], d['ms_per_step'], d['roofline']['achieved'], d['clocks']['sm_mhz'], d['parity'])
P
tail -c 400 $O/c4_n2.err
