#!/bin/bash
# round 2, call g: chord cycle length, lanes, k_control lanes
for v in 1 2 3; do
  echo "== CB_VROUNDS=$v" >> gpurun_out/probe_r2g.log
  CB_VROUNDS=$v timeout 300 python scripts/probe_scale.py 2048:1 16384:4 >> gpurun_out/probe_r2g.log 2>&1
done
echo "== lanes" >> gpurun_out/probe_r2g.log
timeout 300 python scripts/probe_scale.py 16384:3 16384:6 16384:8 8192:2 8192:4 4096:2 >> gpurun_out/probe_r2g.log 2>&1
echo "== CB_CTRL_LANES=8" >> gpurun_out/probe_r2g.log
CB_CTRL_LANES=8 timeout 300 python scripts/probe_scale.py 2048:1 4096:1 >> gpurun_out/probe_r2g.log 2>&1
echo "== CB_CTRL_LANES=32" >> gpurun_out/probe_r2g.log
CB_CTRL_LANES=32 timeout 300 python scripts/probe_scale.py 2048:1 4096:1 16384:4 >> gpurun_out/probe_r2g.log 2>&1
echo "== fixed step" >> gpurun_out/probe_r2g.log
PROBE_FIXED=25e-12 timeout 300 python scripts/probe_scale.py 2048:1 16384:4 >> gpurun_out/probe_r2g.log 2>&1
python -m pytest tests/test_gpu_sweep_api.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_gpu_r2g.log 2>&1
tail -4 gpurun_out/pytest_gpu_r2g.log
