#!/bin/bash
# round 2, call o: control step fused into the tail of k_lu (three rotating list buffers, idle list): parity suite, A/B against CB_FUSE=0
python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu_r2o.log 2>&1
tail -5 gpurun_out/pytest_gpu_r2o.log
run() { echo "== $1 $2 $3" >> gpurun_out/probe_r2o.log; env $1 CB_NVRTC_DEFS=$2 timeout 400 python scripts/probe_scale.py $3 >> gpurun_out/probe_r2o.log 2>&1; }
run "CB_FUSE=1" "" "16384:4 2048:1 4096:1 4096:2 8192:2 8192:4"
run "CB_FUSE=0" "" "16384:4 2048:1 4096:1 4096:2 8192:2 8192:4"
run "CB_FUSE=1" "" "16384:4 2048:1 16384:2 16384:6"
run "CB_FUSE=0" "" "16384:4 2048:1"
cut -c1-150 gpurun_out/probe_r2o.log
