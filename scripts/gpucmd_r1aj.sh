mkdir -p gpurun_out
( time timeout 400 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_r1aj.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu_r1aj.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke_r1aj.log
timeout 240 python bench.py > gpurun_out/bench_r1aj.json 2> gpurun_out/bench_r1aj.err; echo "bench exit $?"; tail -c 600 gpurun_out/bench_r1aj.json
