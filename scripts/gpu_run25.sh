#!/bin/bash
O=gpurun_out/r2t; mkdir -p $O
which compute-sanitizer
timeout 280 compute-sanitizer --tool memcheck --error-exitcode 7 python __graft_entry__.py smoke > $O/memcheck_smoke.txt 2>&1; echo "memcheck rc=$?"
tail -15 $O/memcheck_smoke.txt
