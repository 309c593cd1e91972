#!/bin/bash
# round 2, call d: cache layouts x schedules, launch list of the small-batch case, parity suite
for lay in 3 0 2; do for mixed in 1 0; do
  echo "== VA_CACHE_LAYOUT=$lay CB_MIXED=$mixed" >> gpurun_out/probe_r2d.log
  CB_NVRTC_DEFS=-DVA_CACHE_LAYOUT=$lay CB_MIXED=$mixed timeout 300 python scripts/probe_scale.py 2048:1 16384:4 >> gpurun_out/probe_r2d.log 2>&1
done; done
cat gpurun_out/probe_r2d.log
CB_NOGRAPH=1 CB_MIXED=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 4000 -c 700 --csv \
  --log-file gpurun_out/launches_r2d_b2048_mixed.csv python scripts/probe_scale.py 2048:1 > gpurun_out/ncu_r2d.log 2>&1
python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu_r2d.log 2>&1
tail -8 gpurun_out/pytest_gpu_r2d.log
