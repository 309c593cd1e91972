#!/bin/bash
# round 2, call c: full parity suite (incl. full-span config 3 tests), then lock-step / mixed probes with per-point cache rows
python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu_r2c.log 2>&1
tail -8 gpurun_out/pytest_gpu_r2c.log
for mixed in 0 1; do
  echo "== CB_MIXED=$mixed" >> gpurun_out/probe_r2c.log
  CB_MIXED=$mixed timeout 300 python scripts/probe_scale.py 2048:1 4096:2 16384:4 16384:2 >> gpurun_out/probe_r2c.log 2>&1
done
echo "== fixed CB_MIXED=1" >> gpurun_out/probe_r2c.log
CB_MIXED=1 PROBE_FIXED=25e-12 timeout 300 python scripts/probe_scale.py 2048:1 16384:4 >> gpurun_out/probe_r2c.log 2>&1
cat gpurun_out/probe_r2c.log
