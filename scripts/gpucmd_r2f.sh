#!/bin/bash
# round 2, call f: restored CTA-per-32-points k_lu + lists + graphs: parity suite, in-situ kernel timings, full ncu captures
python -m pytest tests -m gpu -x -q -s --deselect tests/test_gpu_config3_full.py > gpurun_out/pytest_gpu_r2f.log 2>&1
tail -6 gpurun_out/pytest_gpu_r2f.log
echo "== lock-step, graph" >> gpurun_out/probe_r2f.log
timeout 300 python scripts/probe_scale.py 2048:1 4096:1 16384:4 >> gpurun_out/probe_r2f.log 2>&1
echo "== lock-step, CB_TIMING=1 (events around every launch, no graph, one lane)" >> gpurun_out/probe_r2f.log
CB_TIMING=1 timeout 300 python scripts/probe_scale.py 2048:1 16384:1 >> gpurun_out/probe_r2f.log 2>&1
cat gpurun_out/probe_r2f.log
export CB_NOGRAPH=1 CB_LANES=1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_lu|k_control" --launch-skip 3000 -c 6 -f -o gpurun_out/ncu_lu_ctrl_r2f \
   python scripts/probe_scale.py 16384:1 > gpurun_out/ncu_r2f_a.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_lu|k_control" --launch-skip 3000 -c 6 -f -o gpurun_out/ncu_lu_ctrl_b2048_r2f \
   python scripts/probe_scale.py 2048:1 > gpurun_out/ncu_r2f_b.log 2>&1
ls -la gpurun_out/*.ncu-rep
