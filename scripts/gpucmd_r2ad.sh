#!/bin/bash
# round 2, call ad: throughput of the other BASELINE configurations at full size; lane counts at the per-GPU batch sizes of N = 4
timeout 900 python scripts/config_perf.py > gpurun_out/config_perf_r2ad.log 2>&1
cat gpurun_out/config_perf_r2ad.log | cut -c1-300
for c in "4096:2" "4096:4" "4096:3" "2048:2" "8192:4"; do timeout 300 python scripts/probe_scale.py $c >> gpurun_out/probe_r2ad.log 2>&1; done
cut -c1-130 gpurun_out/probe_r2ad.log
