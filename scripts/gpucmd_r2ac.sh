#!/bin/bash
# round 2, call ac: final validation of the committed state -- full gpu suite, smoke, default bench line, reference arm
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r2ac.log 2>&1
tail -3 gpurun_out/pytest_gpu_r2ac.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r2ac.log 2>&1; tail -1 gpurun_out/smoke_r2ac.log
timeout 900 python bench.py > gpurun_out/bench_r2ac.json 2> gpurun_out/bench_r2ac.err; tail -c 200 gpurun_out/bench_r2ac.err; head -c 300 gpurun_out/bench_r2ac.json; echo
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2ac_ref.json 2> gpurun_out/bench_r2ac_ref.err; head -c 300 gpurun_out/bench_r2ac_ref.json; echo
