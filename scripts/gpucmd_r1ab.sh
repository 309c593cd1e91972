mkdir -p gpurun_out
timeout 900 bash scripts/variants_perf.sh scripts/variants_r1ab_run.txt 6e-8 > gpurun_out/variants_r1ab.log 2>&1
grep -v "^==" gpurun_out/variants_r1ab.log
