#!/bin/bash
# round 2, call e: warp-per-point k_lu + per-point dev_out rows: parity suite first, then schedules / launch list
python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu_r2e.log 2>&1
tail -6 gpurun_out/pytest_gpu_r2e.log
for mixed in 1 0; do
  echo "== CB_MIXED=$mixed" >> gpurun_out/probe_r2e.log
  CB_MIXED=$mixed timeout 300 python scripts/probe_scale.py 2048:1 4096:1 8192:2 16384:4 16384:2 16384:1 >> gpurun_out/probe_r2e.log 2>&1
done
cat gpurun_out/probe_r2e.log
CB_NOGRAPH=1 CB_MIXED=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 4000 -c 600 --csv \
  --log-file gpurun_out/launches_r2e_b2048_mixed.csv python scripts/probe_scale.py 2048:1 > gpurun_out/ncu_r2e.log 2>&1
CB_NOGRAPH=1 CB_MIXED=1 CB_LANES=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 4000 -c 600 --csv \
  --log-file gpurun_out/launches_r2e_b16384_mixed.csv python scripts/probe_scale.py 16384:1 > gpurun_out/ncu_r2e2.log 2>&1
