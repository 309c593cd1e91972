#!/bin/bash
# round 2, call ap: chunked row prefetch in the control step (A: default, B: + 3 CTAs per SM for k_control<8>, C: no chunking = before); parity subset first
python -m pytest tests/test_gpu_parity.py tests/test_gpu_lanes.py tests/test_gpu_config3_full.py -m gpu -x -q > gpurun_out/pytest_gpu_r2ap.log 2>&1
tail -3 gpurun_out/pytest_gpu_r2ap.log | cut -c1-250
run() { echo "== $1 $2" >> gpurun_out/probe_r2ap.log; env $1 timeout 300 python scripts/probe_scale.py $2 >> gpurun_out/probe_r2ap.log 2>&1; }
run "CB_X=0" "16384:4 2048:1 4096:2"
run "CB_ENGINE_LIB=scripts/libcedarb200_cu_minb3.so" "16384:4 2048:1 4096:2"
run "CB_ENGINE_LIB=scripts/libcedarb200_cu1.so" "16384:4 2048:1 4096:2"
run "CB_X=0" "16384:4 2048:1"
run "CB_ENGINE_LIB=scripts/libcedarb200_cu1.so" "16384:4 2048:1"
cut -c1-130 gpurun_out/probe_r2ap.log
