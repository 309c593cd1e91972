#!/bin/bash
# round 2, call ai (2 GPUs): the 16 384-point job through ONE multi-device plan of one process (no NCCL), end to end with host buffers
timeout 600 python scripts/multi_plan_perf.py 2 > gpurun_out/multi_plan_r2ai.log 2>&1; cat gpurun_out/multi_plan_r2ai.log | cut -c1-300
timeout 600 python scripts/multi_plan_perf.py 1 >> gpurun_out/multi_plan_r2ai.log 2>&1; tail -1 gpurun_out/multi_plan_r2ai.log | cut -c1-300
