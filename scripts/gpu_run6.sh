#!/bin/bash
O=gpurun_out/r2g; mkdir -p $O
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'vm::' -c 60 --csv --log-file $O/launches_c2_rho0.05.csv \
    python bench.py --steps 2 --warmup 3 --only-main --no-cpu-baseline --no-scaling-baseline --variant clustered --cluster-rho 0.05 > $O/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
ncu --set full --import-source on --clock-control none -k regex:select_rescore_kernel -s 6 -c 1 -f -o $O/select_c2_rho0.05 python bench.py --steps 4 --warmup 3 --only-main --no-cpu-baseline --no-scaling-baseline --variant clustered --cluster-rho 0.05 > $O/ncu_select.log 2>&1; echo "ncu select rc=$?"
