#!/bin/bash
# round 2, call s: cache-ring depth of the value-only eval kernels (ncu: 90 % of their LDS stalls wait for the row copies)
run() { echo "== $1" >> gpurun_out/probe_r2s.log; CB_NVRTC_DEFS=$1 timeout 400 python scripts/probe_scale.py $2 >> gpurun_out/probe_r2s.log 2>&1; }
run "" "16384:4 2048:1"
run "-DVA_AHEAD_V=2,-DVA_STAGES_V=5" "16384:4 2048:1"
run "-DVA_AHEAD_V=3,-DVA_STAGES_V=6" "16384:4 2048:1"
run "-DVA_AHEAD_V=2,-DVA_STAGES_V=6" "16384:4 2048:1"
run "-DVA_AHEAD_V=2,-DVA_STAGES_V=5,-DVA_EVALV_MINBLOCKS=4" "16384:4 2048:1"
run "-DVA_AHEAD=2,-DVA_STAGES=5" "16384:4 2048:1"
run "" "16384:4 2048:1"
cut -c1-135 gpurun_out/probe_r2s.log
