#!/bin/bash
# round 2, call ak: blocked mode of the fused kernel for small batches (no lists, group = 32 consecutive points, control step as the tail of k_lu): full suite + A/B
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r2ak.log 2>&1
tail -12 gpurun_out/pytest_gpu_r2ak.log | cut -c1-250
run() { echo "== $1 $2" >> gpurun_out/probe_r2ak.log; env $1 timeout 400 python scripts/probe_scale.py $2 >> gpurun_out/probe_r2ak.log 2>&1; }
run "CB_X=0" "2048:1 4096:2 4096:1 1024:1 16384:4"
run "CB_BLOCKED=0" "2048:1 4096:2 4096:1 1024:1"
run "CB_X=0" "2048:1 4096:2"
run "CB_BLOCKED=0" "2048:1 4096:2"
cut -c1-135 gpurun_out/probe_r2ak.log
