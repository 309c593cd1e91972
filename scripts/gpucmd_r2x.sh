#!/bin/bash
# round 2, call x: new GPU tests (retry ladder, switch branch, native front-end extensions) + the full suite
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r2x.log 2>&1
tail -6 gpurun_out/pytest_gpu_r2x.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r2x.log 2>&1; tail -2 gpurun_out/smoke_r2x.log
