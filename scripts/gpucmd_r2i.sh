#!/bin/bash
# round 2, call i: full parity suite, L2 evict-first hint A/B, first full bench line
python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu_r2i.log 2>&1
tail -6 gpurun_out/pytest_gpu_r2i.log
for rep in 1 2; do
  echo "== default (rep $rep)" >> gpurun_out/probe_r2i.log
  timeout 300 python scripts/probe_scale.py 16384:4 >> gpurun_out/probe_r2i.log 2>&1
  echo "== VA_CACHE_HINT=1 (rep $rep)" >> gpurun_out/probe_r2i.log
  CB_NVRTC_DEFS=-DVA_CACHE_HINT=1 timeout 300 python scripts/probe_scale.py 16384:4 >> gpurun_out/probe_r2i.log 2>&1
done
echo "== small" >> gpurun_out/probe_r2i.log
timeout 300 python scripts/probe_scale.py 2048:1 4096:2 >> gpurun_out/probe_r2i.log 2>&1
cat gpurun_out/probe_r2i.log | cut -c1-220
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r2i.json 2> gpurun_out/bench_r2i.err
tail -c 600 gpurun_out/bench_r2i.err; head -c 1500 gpurun_out/bench_r2i.json
