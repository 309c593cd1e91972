mkdir -p gpurun_out
timeout 900 bash scripts/variants_perf.sh scripts/variants_r1x_run.txt 6e-8 > gpurun_out/variants_r1x.log 2>&1
grep -v "^==" gpurun_out/variants_r1x.log
timeout 600 python -m pytest tests/test_gpu_lanes.py tests/test_gpu_parity.py -q -m gpu 2>&1 | tail -3
