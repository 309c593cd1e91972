#!/bin/bash
# round 2, call aj: the gpu suite and smoke after pruning the cubin cache of stale experiment variants (what is missing is compiled by NVRTC on the box)
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r2aj.log 2>&1
tail -3 gpurun_out/pytest_gpu_r2aj.log | cut -c1-250
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
