mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
(time timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu_r1u.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_r1u.log
timeout 900 bash scripts/variants_perf.sh scripts/variants_r1u_run.txt 6e-8 > gpurun_out/variants_r1u.log 2>&1
tail -3 gpurun_out/pytest_gpu_r1u.log; cat gpurun_out/variants_r1u.log
