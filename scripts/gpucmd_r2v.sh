#!/bin/bash
# round 2, call v: why is the control step as the tail of k_lu slow?  full ncu capture (source page) of the fused kernel at 2 048 points
export CB_NOGRAPH=1 CB_LANES=1 CB_FUSE=1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_lu" --launch-skip 1500 -c 3 -f -o gpurun_out/ncu_lu_fused_b2048_r2v \
   python scripts/probe_scale.py 2048:1 > gpurun_out/ncu_r2v.log 2>&1
tail -3 gpurun_out/ncu_r2v.log
ls -la gpurun_out/*r2v*
