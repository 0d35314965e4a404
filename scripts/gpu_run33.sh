#!/bin/bash
# round 2, final 1-GPU validation + refreshed captures
O=gpurun_out/r2y; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.txt
tail -5 $O/pytest_gpu.txt
timeout 300 python __graft_entry__.py smoke > $O/smoke.txt 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'scan_tc|select_rescore|normalize_queries|exact_scan' -c 60 --csv --log-file $O/launches_c2.csv \
    python bench.py --steps 2 --warmup 3 --only-main --no-cpu-baseline --no-scaling-baseline > $O/ncu_launch_c2.log 2>&1; echo "ncu launches c2 rc=$?"
ncu --set full --import-source on --clock-control none -k regex:scan_tc_kernel -s 6 -c 1 -f -o $O/scan_c2_f32 python bench.py --steps 4 --warmup 3 --only-main --no-cpu-baseline --no-scaling-baseline > $O/ncu_scan_c2.log 2>&1; echo "ncu scan c2 rc=$?"
ncu --set full --import-source on --clock-control none -k regex:select_rescore_kernel -s 6 -c 1 -f -o $O/select_c2_f32 python bench.py --steps 4 --warmup 3 --only-main --no-cpu-baseline --no-scaling-baseline > $O/ncu_select_c2.log 2>&1; echo "ncu select c2 rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2y/bench_n1.json').read().strip().splitlines()[-1])
print('main', round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), d['parity']['ok'], d['roofline']['frac'], d['roofline']['step_frac'])
for k in ('clustered','dedup','streaming','streaming_bf16'):
    r=d[k]; print(k, r['value'], r['ms_per_step'], (r.get('parity') or {}).get('ok'), r.get('latency_ms'))
for k,r in d['binary64_store'].items(): print(k, r['value'], r['parity']['ok'])
print(json.loads(open('gpurun_out/r2y/bench_ref.json').read().strip().splitlines()[-1])['value'])
P
tail -3 $O/smoke.txt
