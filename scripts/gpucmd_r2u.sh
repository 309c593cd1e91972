#!/bin/bash
# round 2, call u: control step with a small instruction footprint (out-of-line IEEE division, no unrolling, out-of-line waveform / interpolation helpers), stand-alone and as the tail of k_lu
python -m pytest tests/test_gpu_parity.py tests/test_gpu_lanes.py -m gpu -x -q > gpurun_out/pytest_gpu_r2u.log 2>&1
tail -3 gpurun_out/pytest_gpu_r2u.log
run() { echo "== $1 $2" >> gpurun_out/probe_r2u.log; env $1 timeout 400 python scripts/probe_scale.py $2 >> gpurun_out/probe_r2u.log 2>&1; }
run "CB_X=0" "16384:4 2048:1 4096:2"
run "CB_ENGINE_LIB=scripts/libcedarb200_bigctrl.so" "16384:4 2048:1 4096:2"
run "CB_FUSE=1" "16384:4 2048:1 4096:2"
run "CB_X=0" "16384:4 2048:1 4096:2"
run "CB_ENGINE_LIB=scripts/libcedarb200_bigctrl.so" "16384:4 2048:1"
cut -c1-135 gpurun_out/probe_r2u.log
