mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q) > gpurun_out/pytest_gpu_r1ae.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_r1ae.log
tail -25 gpurun_out/pytest_gpu_r1ae.log
(time timeout 900 python bench.py --steps 2 --warmup 3) > gpurun_out/bench_r1ae.json 2> gpurun_out/bench_r1ae.err
tail -c 400 gpurun_out/bench_r1ae.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
