#!/bin/bash
# round 2, call y: full suite again (switch-branch test tolerance), linear stamp values of k_lu staged in shared memory (A/B: CB_LU_NOLIN=1)
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r2y.log 2>&1
tail -4 gpurun_out/pytest_gpu_r2y.log
run() { echo "== $1 $2" >> gpurun_out/probe_r2y.log; env $1 timeout 400 python scripts/probe_scale.py $2 >> gpurun_out/probe_r2y.log 2>&1; }
run "CB_X=0" "16384:4 2048:1"
run "CB_LU_NOLIN=1" "16384:4 2048:1"
run "CB_X=0" "16384:4 2048:1"
run "CB_LU_NOLIN=1" "16384:4 2048:1"
cut -c1-135 gpurun_out/probe_r2y.log
