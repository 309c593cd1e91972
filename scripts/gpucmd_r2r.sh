#!/bin/bash
# round 2, call r: k_lu L2 prefetch (charge-update rows, next group's rows) and row preloading, A/B against the variants; stream priorities
python -m pytest tests/test_gpu_parity.py tests/test_gpu_lanes.py -m gpu -x -q > gpurun_out/pytest_gpu_r2r.log 2>&1
tail -3 gpurun_out/pytest_gpu_r2r.log
run() { echo "== $1 $2" >> gpurun_out/probe_r2r.log; env $1 timeout 400 python scripts/probe_scale.py $2 >> gpurun_out/probe_r2r.log 2>&1; }
run "CB_X=0" "16384:4 2048:1 4096:1"
for v in pf0 pf1 pf2; do run "CB_ENGINE_LIB=scripts/libcedarb200_$v.so" "16384:4 2048:1"; done
run "CB_PRIO=1" "16384:4 16384:6 8192:4"
run "CB_X=0" "16384:4 2048:1 8192:4"
cut -c1-150 gpurun_out/probe_r2r.log
