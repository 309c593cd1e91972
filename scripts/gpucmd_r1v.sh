mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q) > gpurun_out/pytest_gpu_r1v.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_r1v.log
timeout 900 bash scripts/variants_perf.sh scripts/variants_r1v_run.txt 6e-8 > gpurun_out/variants_r1v.log 2>&1
tail -5 gpurun_out/pytest_gpu_r1v.log; grep -v "^==" gpurun_out/variants_r1v.log
