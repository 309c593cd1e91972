#!/bin/bash
# round 2, call m: uniform cache slots + atomic reductions in k_lu (parity suite), register variants of the eval kernels at small batches
python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu_r2m.log 2>&1
tail -5 gpurun_out/pytest_gpu_r2m.log
run() { echo "== $1 $2" >> gpurun_out/probe_r2m.log; env $1 timeout 400 python scripts/probe_scale.py $2 >> gpurun_out/probe_r2m.log 2>&1; }
run "CB_X=0" "16384:4 2048:1 4096:1 4096:2"
run "CB_MAXREG=255" "2048:1 4096:1"
run "CB_MAXREG=168" "2048:1 4096:1"
run "CB_MAXREG=128" "2048:1 4096:1 16384:4"
run "CB_X=0" "16384:4 2048:1"
cut -c1-150 gpurun_out/probe_r2m.log
