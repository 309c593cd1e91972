#!/bin/bash
# round 2, call ah: full gpu suite with the repair pass as an opt-in kernel instantiation; the normal path's speed
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r2ah.log 2>&1
tail -3 gpurun_out/pytest_gpu_r2ah.log | cut -c1-250
timeout 400 python scripts/probe_scale.py 16384:4 2048:1 > gpurun_out/probe_r2ah.log 2>&1; cut -c1-135 gpurun_out/probe_r2ah.log
