#!/bin/bash
# round 2, call k (2 GPUs): multi-device plan test, strong-scaling bench at N = 2 under torchrun
python -m pytest tests/test_gpu_lanes.py -m gpu -x -q > gpurun_out/pytest_gpu_r2k_2gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu_r2k_2gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 \
   > gpurun_out/bench_r2k_n2.json 2> gpurun_out/bench_r2k_n2.err
tail -c 400 gpurun_out/bench_r2k_n2.err; head -c 1200 gpurun_out/bench_r2k_n2.json
