#!/bin/bash
# round 2, call j: profiles -- launch list of the bench command, full ncu captures of the five kernels, compute-sanitizer
export CB_NOGRAPH=1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 12000 -c 600 --csv --log-file gpurun_out/launches_r2j.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra-legs > gpurun_out/ncu_r2j_bench.log 2>&1
export CB_LANES=1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_eval_bsimcmg107_nmos|k_evalv_bsimcmg107_nmos|k_lu|k_control" \
   --launch-skip 3000 -c 10 -f -o gpurun_out/ncu_kernels_r2j python scripts/probe_scale.py 16384:1 > gpurun_out/ncu_r2j_a.log 2>&1
unset CB_NOGRAPH CB_LANES
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_memcheck_r2j.log 2>&1
tail -4 gpurun_out/sanitizer_memcheck_r2j.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_racecheck_r2j.log 2>&1
tail -4 gpurun_out/sanitizer_racecheck_r2j.log
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_r2j.csv
