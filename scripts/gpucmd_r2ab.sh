#!/bin/bash
# round 2, call ab (4 GPUs): the bench line under torchrun at N = 4 (4 096 points per GPU)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 4 --steps 3 --warmup 3 \
   > gpurun_out/bench_r2ab_n4.json 2> gpurun_out/bench_r2ab_n4.err
tail -c 300 gpurun_out/bench_r2ab_n4.err; head -c 400 gpurun_out/bench_r2ab_n4.json; echo
