#!/bin/bash
# round 2, call 9: warp-cooperative own-k-th in the bound service, PDL inside the captured graph; ncu of the scan on a tight-cluster store
O=gpurun_out/r2j; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.txt
tail -6 $O/pytest_gpu.txt
timeout 300 python bench.py --steps 20 --warmup 5 --only-main --no-scaling-baseline --no-cpu-baseline > $O/bench_c2.json 2> $O/bench_c2.err; echo "c2 rc=$?"
for rho in 0.05 0.2; do
timeout 300 python bench.py --steps 20 --warmup 5 --only-main --no-scaling-baseline --variant clustered --cluster-rho $rho --no-cpu-baseline > $O/bench_c2_clustered_rho$rho.json 2> $O/bench_c2_clustered_rho$rho.err; echo "c2 clustered rho=$rho rc=$?"
done
timeout 300 python bench.py --steps 20 --warmup 5 --only-main --config c3 --rows 12500000 --no-cpu-baseline > $O/bench_c3shard.json 2> $O/bench_c3shard.err; echo "c3 shard rc=$?"
timeout 300 python bench.py --steps 20 --warmup 5 --only-main --config c3 --rows 12500000 --no-cpu-baseline --variant clustered > $O/bench_c3shard_clustered.json 2> $O/bench_c3shard_clustered.err; echo "c3 shard clustered rc=$?"
timeout 300 python bench.py --steps 20 --warmup 5 --only-main --no-scaling-baseline --no-cpu-baseline --store-dtype bf16 > $O/bench_c2_bf16.json 2> $O/bench_c2_bf16.err; echo "c2 bf16 rc=$?"
timeout 300 python bench.py --steps 30 --warmup 5 --only-main --config c1 --no-cpu-baseline > $O/bench_c1.json 2> $O/bench_c1.err; echo "c1 rc=$?"
ncu --set full --import-source on --clock-control none -k regex:scan_tc_kernel -s 6 -c 1 -f -o $O/scan_c2_rho0.05 python bench.py --steps 4 --warmup 3 --only-main --no-cpu-baseline --no-scaling-baseline --variant clustered --cluster-rho 0.05 > $O/ncu_scan.log 2>&1; echo "ncu scan rc=$?"
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2j/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']; print(f.split('/')[-1], round(d['value']), round(d['ms_per_step'],4), 'scan', round(r['kernel_ms'],4), 'e2e', round(d['e2e']['value']), d['parity']['ok'])
    except Exception as e: print(f,'ERR',e)
P
tail -c 300 $O/*.err
