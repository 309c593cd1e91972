#!/bin/bash
# round 2, call aa (8 GPUs): the bench line under torchrun at N = 8 -- BASELINE config 3 as written (16 384 points total, 2 048 per GPU) + weak leg
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 3 --warmup 3 \
   > gpurun_out/bench_r2aa_n8.json 2> gpurun_out/bench_r2aa_n8.err
tail -c 300 gpurun_out/bench_r2aa_n8.err; head -c 900 gpurun_out/bench_r2aa_n8.json; echo
