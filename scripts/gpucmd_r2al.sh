#!/bin/bash
# round 2, call al: full suite with blocked mode on for small batches (repair pass keeps the list-based kernel)
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r2al.log 2>&1
tail -6 gpurun_out/pytest_gpu_r2al.log | cut -c1-250
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
