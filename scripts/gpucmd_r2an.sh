#!/bin/bash
# round 2, call an (4 GPUs): strong / weak scaling with blocked mode on at 4 096 points per GPU
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 4 --steps 3 --warmup 3 \
   > gpurun_out/bench_r2an_n4.json 2> gpurun_out/bench_r2an_n4.err
tail -c 200 gpurun_out/bench_r2an_n4.err; head -c 300 gpurun_out/bench_r2an_n4.json; echo
