#!/bin/bash
# round 2, call w (2 GPUs): multi-device plan tests and the bench line under torchrun (strong scaling: 8 192 points per GPU)
python -m pytest tests/test_gpu_lanes.py -m gpu -x -q > gpurun_out/pytest_gpu_r2w_2gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu_r2w_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 2 --warmup 3 \
   > gpurun_out/bench_r2w_n2.json 2> gpurun_out/bench_r2w_n2.err
tail -c 300 gpurun_out/bench_r2w_n2.err; head -c 700 gpurun_out/bench_r2w_n2.json; echo
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 \
   > gpurun_out/bench_r2w_ref_n2.json 2> gpurun_out/bench_r2w_ref_n2.err
head -c 500 gpurun_out/bench_r2w_ref_n2.json; echo
