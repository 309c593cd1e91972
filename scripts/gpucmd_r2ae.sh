#!/bin/bash
# round 2, call ae: repair pass of k_lu (flagged points re-solved with partial pivoting): its test, the parity suite, and the normal path's speed
python -m pytest tests/test_gpu_parity.py tests/test_gpu_lanes.py tests/test_gpu_sweep_api.py -m gpu -x -q > gpurun_out/pytest_gpu_r2ae.log 2>&1
tail -15 gpurun_out/pytest_gpu_r2ae.log | cut -c1-250
run() { echo "== $1 $2" >> gpurun_out/probe_r2ae.log; env $1 timeout 400 python scripts/probe_scale.py $2 >> gpurun_out/probe_r2ae.log 2>&1; }
run "CB_X=0" "16384:4 2048:1"
run "CB_X=0" "16384:4 2048:1"
cut -c1-135 gpurun_out/probe_r2ae.log
