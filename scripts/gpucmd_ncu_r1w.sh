mkdir -p gpurun_out
export RELTOL=1e-4 NR_RELTOL=1e-5 NR_VABSTOL=1e-7 NR_IABSTOL=1e-13 NR_RATE_TEST=1 VALUE_ROUNDS=2 CB_LANES=1 CB_EVAL_FORK=0 CB_MAX_ROUNDS=400
run() { # name regex skip
  timeout 600 ncu --set full --clock-control none -k "regex:$2" -s $3 -c 1 -f -o gpurun_out/ncu_$1_r1w python scripts/first_perf.py 16384 adaptive 6e-8 > gpurun_out/ncu_$1_r1w.log 2>&1
  echo "$1 rc=$?"
}
run eval   '^k_eval_bsimcmg107_nmos' 60
run evalv  '^k_evalv_bsimcmg107_nmos' 60
run lu     'k_lu<false|k_lu<\(bool\)0' 60
run lus    'k_lu<true|k_lu<\(bool\)1' 60
run ctrl   '^k_control' 150
ls -la gpurun_out/*.ncu-rep
