mkdir -p gpurun_out
timeout 900 bash scripts/variants_perf.sh scripts/variants_r1w_run.txt 6e-8 > gpurun_out/variants_r1w.log 2>&1
grep -v "^==" gpurun_out/variants_r1w.log
(time timeout 900 python bench.py --steps 2 --warmup 3) > gpurun_out/bench_r1w.json 2> gpurun_out/bench_r1w.err
tail -c 1500 gpurun_out/bench_r1w.json
