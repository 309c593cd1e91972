#!/bin/bash
# round 2, call af: repair-pass test with the oracle leg; same-box A/B of the normal path against the previous build
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pivot or flagged" > gpurun_out/pytest_gpu_r2af.log 2>&1
tail -3 gpurun_out/pytest_gpu_r2af.log | cut -c1-250
run() { echo "== $1 $2" >> gpurun_out/probe_r2af.log; env $1 timeout 400 python scripts/probe_scale.py $2 >> gpurun_out/probe_r2af.log 2>&1; }
for rep in 1 2 3; do
run "CB_X=0" "16384:4 2048:1"
run "CB_ENGINE_LIB=scripts/libcedarb200_prev.so" "16384:4 2048:1"
done
cut -c1-135 gpurun_out/probe_r2af.log
