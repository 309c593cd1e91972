#!/bin/bash
# round 2, call h: full parity suite (t0 re-initialisation, direct sensitivities, config 3 full span), A/B of the k_lu change
python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu_r2h.log 2>&1
tail -6 gpurun_out/pytest_gpu_r2h.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit,temperature.gpu --format=csv >> gpurun_out/probe_r2h.log
for rep in 1 2; do
  echo "== default lib (rep $rep)" >> gpurun_out/probe_r2h.log
  timeout 300 python scripts/probe_scale.py 16384:4 >> gpurun_out/probe_r2h.log 2>&1
  echo "== no-unroll lib (rep $rep)" >> gpurun_out/probe_r2h.log
  CB_ENGINE_LIB=$PWD/scripts/libcedarb200_nounroll.so timeout 300 python scripts/probe_scale.py 16384:4 >> gpurun_out/probe_r2h.log 2>&1
done
echo "== small batches" >> gpurun_out/probe_r2h.log
CB_LANES=2 timeout 300 python scripts/probe_scale.py 2048:2 4096:2 >> gpurun_out/probe_r2h.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit,temperature.gpu --format=csv >> gpurun_out/probe_r2h.log
cat gpurun_out/probe_r2h.log
