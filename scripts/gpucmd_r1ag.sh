mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q) > gpurun_out/pytest_gpu_r1ag.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_r1ag.log
tail -25 gpurun_out/pytest_gpu_r1ag.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
