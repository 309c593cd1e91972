#!/bin/bash
# round 2, call l: layout of the value-only cache (its lock-step launches have gaps), value-only occupancy; single-lane launch list
run() { echo "== $1" >> gpurun_out/probe_r2l.log; CB_NVRTC_DEFS=$1 timeout 300 python scripts/probe_scale.py 16384:4 2048:1 >> gpurun_out/probe_r2l.log 2>&1; }
run ""
run "-DVA_CACHE_LAYOUT_V=3"
run "-DVA_CACHE_LAYOUT_V=2"
run ""
run "-DVA_EVALV_MINBLOCKS=6"
run "-DVA_CACHE_LAYOUT_V=3,-DVA_EVALV_MINBLOCKS=6"
cut -c1-200 gpurun_out/probe_r2l.log
CB_NOGRAPH=1 CB_LANES=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 4000 -c 600 --csv --log-file gpurun_out/launches_r2l_1lane.csv \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra-legs > gpurun_out/ncu_r2l_bench.log 2>&1
