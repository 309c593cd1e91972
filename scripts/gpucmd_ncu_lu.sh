mkdir -p gpurun_out
export RELTOL=1e-4 NR_RELTOL=1e-5 NR_VABSTOL=1e-7 NR_IABSTOL=1e-13 NR_RATE_TEST=1 VALUE_ROUNDS=2 CB_LANES=1 CB_EVAL_FORK=0 CB_MAX_ROUNDS=400
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_lu -s 90 -c 3 -f -o gpurun_out/ncu_lu_r1ah python scripts/first_perf.py 16384 adaptive 6e-8 > gpurun_out/ncu_lu_r1ah.log 2>&1
echo "lu rc=$?"; ls -la gpurun_out/ncu_lu_r1ah.ncu-rep
