#!/bin/bash
# round 2, call q: profiles of the current kernels -- single-lane launch list, full ncu captures (source page) of the five kernels at 16 384 and k_lu / k_control at 2 048 points; new $abstime GPU test
python -m pytest tests/test_gpu_sweep_api.py -m gpu -x -q -k abstime > gpurun_out/pytest_gpu_r2q.log 2>&1
tail -3 gpurun_out/pytest_gpu_r2q.log
export CB_NOGRAPH=1 CB_LANES=1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 4000 -c 600 --csv --log-file gpurun_out/launches_r2q_1lane.csv \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra-legs > gpurun_out/ncu_r2q_bench.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_eval_bsimcmg107_nmos|k_evalv_bsimcmg107_nmos|k_lu|k_control" \
   --launch-skip 3000 -c 10 -f -o gpurun_out/ncu_kernels_r2q python scripts/probe_scale.py 16384:1 > gpurun_out/ncu_r2q_a.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_lu|k_control" \
   --launch-skip 1200 -c 6 -f -o gpurun_out/ncu_lu_ctrl_b2048_r2q python scripts/probe_scale.py 2048:1 > gpurun_out/ncu_r2q_b.log 2>&1
ls -la gpurun_out/*r2q*
