mkdir -p gpurun_out
timeout 900 bash scripts/variants_perf.sh scripts/variants_r1ad_run.txt 6e-8 > gpurun_out/variants_r1ad.log 2>&1
grep -v "^==" gpurun_out/variants_r1ad.log
(time timeout 1500 python -m pytest tests -m gpu -q) > gpurun_out/pytest_gpu_r1ad.log 2>&1
tail -15 gpurun_out/pytest_gpu_r1ad.log
