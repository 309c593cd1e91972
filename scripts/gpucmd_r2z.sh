#!/bin/bash
# round 2, call z: schedule knobs after the k_lu speed-up -- value-only rounds per full round, poll window
run() { echo "== $1 $2" >> gpurun_out/probe_r2z.log; env $1 timeout 400 python scripts/probe_scale.py $2 >> gpurun_out/probe_r2z.log 2>&1; }
run "CB_X=0" "16384:4 2048:1"
run "CB_VROUNDS=1" "16384:4 2048:1"
run "CB_VROUNDS=3" "16384:4 2048:1"
run "CB_VROUNDS=4" "16384:4 2048:1"
run "CB_POLL=54" "16384:4 2048:1"
run "CB_POLL=108" "2048:1"
run "CB_X=0" "16384:4 2048:1"
cut -c1-230 gpurun_out/probe_r2z.log
