mkdir -p gpurun_out
timeout 900 bash scripts/variants_perf.sh scripts/variants_r1af_run.txt 6e-8 > gpurun_out/variants_r1af.log 2>&1
grep -v "^==" gpurun_out/variants_r1af.log
CB_ENGINE_LIB=cedarsim.jl_b200/csrc/lib_p16w16.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_lanes.py -q -m gpu 2>&1 | tail -4
