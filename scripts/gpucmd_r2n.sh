#!/bin/bash
# round 2, call n: k_lu with its index tables staged in shared memory (parity, A/B against CB_LU_NOSTAGE), register variants of the eval kernels at small batches
python -m pytest tests/test_gpu_parity.py tests/test_gpu_lanes.py tests/test_gpu_sweep_api.py -m gpu -x -q > gpurun_out/pytest_gpu_r2n.log 2>&1
tail -5 gpurun_out/pytest_gpu_r2n.log
run() { echo "== $1 $2 $3" >> gpurun_out/probe_r2n.log; env $1 CB_NVRTC_DEFS=$2 timeout 400 python scripts/probe_scale.py $3 >> gpurun_out/probe_r2n.log 2>&1; }
run "CB_X=0" "" "16384:4 2048:1 4096:1"
run "CB_LU_NOSTAGE=1" "" "16384:4 2048:1"
run "CB_X=0" "-DVA_EVAL_MINBLOCKS=2" "2048:1 4096:1"
run "CB_X=0" "-DVA_EVAL_MINBLOCKS=3" "2048:1 4096:1"
run "CB_X=0" "-DVA_EVAL_MINBLOCKS=2,-DVA_EVALV_MINBLOCKS=3" "2048:1"
run "CB_X=0" "" "16384:4 2048:1"
cut -c1-150 gpurun_out/probe_r2n.log
