#!/bin/bash
run() { echo "== $*"; env "$@" CB_TIMING=1 python scripts/first_perf.py 16384 adaptive 6e-8 2>&1 | tail -1 | sed 's/.*rounds/rounds/'; }
run A=1
run CB_NVRTC_DEFS=-DVA_EVAL_MINBLOCKS=10
run CB_NVRTC_DEFS=-DVA_EVAL_MINBLOCKS=12
run CB_NVRTC_DEFS=-DVA_EVAL_MINBLOCKS=6
