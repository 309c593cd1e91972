#!/bin/bash
# round 2, call t: full parity suite (native front end, $abstime, C harness from deck text), bench line, compute-sanitizer on the new k_lu
python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu_r2t.log 2>&1
tail -5 gpurun_out/pytest_gpu_r2t.log
timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r2t.json 2> gpurun_out/bench_r2t.err
tail -c 400 gpurun_out/bench_r2t.err; head -c 1200 gpurun_out/bench_r2t.json; echo
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_memcheck_r2t.log 2>&1
tail -4 gpurun_out/sanitizer_memcheck_r2t.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_racecheck_r2t.log 2>&1
tail -4 gpurun_out/sanitizer_racecheck_r2t.log
