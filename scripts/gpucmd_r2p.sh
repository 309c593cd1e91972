#!/bin/bash
# round 2, call p: parity suite with the three list buffers (control step not fused), gather depth of k_lu and k_control occupancy variants
python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu_r2p.log 2>&1
tail -4 gpurun_out/pytest_gpu_r2p.log
run() { echo "== $1 $2" >> gpurun_out/probe_r2p.log; env $1 timeout 400 python scripts/probe_scale.py $2 >> gpurun_out/probe_r2p.log 2>&1; }
run "CB_X=0" "16384:4 2048:1"
for v in gu4 gu12 gu16 ctrl3 ctrl2; do run "CB_ENGINE_LIB=scripts/libcedarb200_$v.so" "16384:4 2048:1"; done
run "CB_X=0" "16384:4 2048:1"
cut -c1-150 gpurun_out/probe_r2p.log
