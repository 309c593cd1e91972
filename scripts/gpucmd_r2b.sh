#!/bin/bash
# round 2, call b: parity suite on the list-based rounds, then lock-step / mixed x graph / no-graph probes
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r2b.log 2>&1
tail -5 gpurun_out/pytest_gpu_r2b.log
for mixed in 0 1; do for nog in 1 0; do
  echo "== CB_MIXED=$mixed CB_NOGRAPH=$nog" >> gpurun_out/probe_r2b.log
  CB_MIXED=$mixed CB_NOGRAPH=$nog timeout 300 python scripts/probe_scale.py 2048:1 16384:4 >> gpurun_out/probe_r2b.log 2>&1
done; done
echo "== fixed CB_MIXED=1" >> gpurun_out/probe_r2b.log
CB_MIXED=1 PROBE_FIXED=25e-12 timeout 300 python scripts/probe_scale.py 2048:1 16384:4 >> gpurun_out/probe_r2b.log 2>&1
cat gpurun_out/probe_r2b.log
