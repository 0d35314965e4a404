#!/bin/bash
O=gpurun_out/r2r; mkdir -p $O
timeout 900 python -m pytest tests/test_pairs_gpu.py tests/test_fullsize_gpu.py tests/test_shards_one_gpu.py -x -q > $O/pytest_pairs.txt 2>&1; echo "pytest rc=$?" >> $O/pytest_pairs.txt
tail -4 $O/pytest_pairs.txt
timeout 300 python bench.py --config c4 --steps 3 --warmup 1 > $O/c4.json 2> $O/c4.err
timeout 300 python bench.py --config c4 --rows 262144 --steps 3 --warmup 1 --no-cpu-baseline > $O/c4_262k.json 2> $O/c4_262k.err
timeout 300 python bench.py --config c4 --rows 500000 --steps 3 --warmup 1 --no-cpu-baseline > $O/c4_500k.json 2> $O/c4_500k.err
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'pairs' -c 12 --csv --log-file $O/launches_c4_262k.csv \
    python bench.py --config c4 --rows 262144 --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_launch_c4.log 2>&1; echo "ncu launches c4 rc=$?"
ncu --set full --clock-control none -k regex:pairs_tc2 -c 1 -f -o $O/pairs_c4_262k python bench.py --config c4 --rows 262144 --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_pairs.log 2>&1; echo "ncu pairs rc=$?"
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2r/c4*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f.split('/')[-1], '%.4g'%d['value'], round(d['ms_per_step'],3), 'TF', round(d['roofline']['achieved'],1), d['roofline']['frac'], d['clocks']['sm_mhz'], d['parity'], d['e2e'])
    except Exception as e: print(f,'ERR',e)
P
tail -c 300 $O/*.err
