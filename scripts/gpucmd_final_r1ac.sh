mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q) > gpurun_out/pytest_gpu_r1ac.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_r1ac.log
tail -4 gpurun_out/pytest_gpu_r1ac.log
(time timeout 900 python bench.py --steps 2 --warmup 3) > gpurun_out/bench_r1ac.json 2> gpurun_out/bench_r1ac.err
tail -c 600 gpurun_out/bench_r1ac.json
export RELTOL=1e-4 NR_RELTOL=1e-5 NR_VABSTOL=1e-7 NR_IABSTOL=1e-13 NR_RATE_TEST=1 VALUE_ROUNDS=2 CB_LANES=1 CB_EVAL_FORK=0 CB_MAX_ROUNDS=400
run() { # name regex skip
  timeout 600 ncu --set full --clock-control none -k "regex:$2" -s $3 -c 1 -f -o gpurun_out/ncu_$1_r1ac python scripts/first_perf.py 16384 adaptive 6e-8 > gpurun_out/ncu_$1_r1ac.log 2>&1
  echo "$1 rc=$?"
}
run eval   '^k_eval_bsimcmg107_nmos' 60
run evalv  '^k_evalv_bsimcmg107_nmos' 100
run lu     'k_lu.*0.*64' 60
run lus    'k_lu.*1.*64' 100
run ctrl   'k_control' 150
unset CB_LANES CB_EVAL_FORK CB_MAX_ROUNDS
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 4000 -c 600 --csv --log-file gpurun_out/launches_r1ac.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/launches_r1ac.log 2>&1
echo "launch list rc=$?"
ls -la gpurun_out/
