#!/bin/bash
# N-GPU evidence run (N = 2, 4, 8), from the repo root on the GPU box:
#     gpurun --gpus N --timeout 1500 -- 'scripts/gpu_evidence_multi.sh N gpurun_out/<tag>'
# The 2-GPU parity tests, then the contract's launch line for the complete bench line (C3 row-sharded + the dedup /
# streaming / clustered sub-records with their in-run parity) and the reference arm.
N=${1:-2}; O=${2:-gpurun_out/evidence_n$N}; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531"
timeout 900 python -m pytest tests/test_sharded_gpu.py tests/test_pairs_gpu.py -m gpu -x -q > $O/pytest_sharded.txt 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_sharded.txt
tail -4 $O/pytest_sharded.txt
timeout 1200 $TR bench.py --gpus $N > $O/bench_n$N.json 2> $O/bench_n$N.err; echo "bench rc=$?"
timeout 600 $TR bench.py --impl reference --gpus $N --steps 2 --warmup 1 > $O/bench_ref_n$N.json 2> $O/bench_ref_n$N.err; echo "reference arm rc=$?"
tail -c 400 $O/bench_n$N.err
