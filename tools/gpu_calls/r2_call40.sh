#!/bin/bash
# round 2, GPU call 40: debugging the split windows (tensor-core path vs CUDA-core path, L = 129 ... 300)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python tools/experiments/exp_long.py > gpurun_out/r2_x40_default.log 2>&1
LS=130,200 RBN_PDL=0 timeout 300 python tools/experiments/exp_long.py > gpurun_out/r2_x40_pdl0.log 2>&1
LS=130,200 RBN_Y_CACHE_GB=0 timeout 300 python tools/experiments/exp_long.py > gpurun_out/r2_x40_nocache.log 2>&1
tail -20 gpurun_out/r2_x40_*.log
