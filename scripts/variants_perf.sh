#!/bin/bash
run() { echo "== $*"; env "$@" CB_TIMING=1 python scripts/first_perf.py 16384 adaptive 6e-8 2>&1 | tail -1 | sed 's/.*rounds/rounds/'; }
run A=1
run CB_NVRTC_DEFS=-DVA_EVAL_MINBLOCKS=3
run CB_NVRTC_DEFS=-DVA_EVAL_MINBLOCKS=4
run CB_NVRTC_DEFS=-DVA_EVAL_MINBLOCKS=6
run "CB_NVRTC_DEFS=-DVA_EVAL_THREADS=64 -DVA_EVAL_MINBLOCKS=6" CB_EVAL_THREADS=64
