#!/bin/bash
# round 2, call ao: the default bench command on the final commit
timeout 900 python bench.py > gpurun_out/bench_r2ao.json 2> gpurun_out/bench_r2ao.err; tail -c 200 gpurun_out/bench_r2ao.err; head -c 300 gpurun_out/bench_r2ao.json; echo
