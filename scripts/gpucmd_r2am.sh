#!/bin/bash
# round 2, call am (8 GPUs): strong / weak scaling with blocked mode on at 2 048 points per GPU
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 3 --warmup 3 \
   > gpurun_out/bench_r2am_n8.json 2> gpurun_out/bench_r2am_n8.err
tail -c 200 gpurun_out/bench_r2am_n8.err; head -c 300 gpurun_out/bench_r2am_n8.json; echo
