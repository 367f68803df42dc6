#!/bin/bash
# round 2, GPU call 41: split windows, finer lengths; with and without the split-sum cache
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
LS=200,225,241,242,249,256,257 timeout 300 python tools/experiments/exp_long.py > gpurun_out/r2_x41_default.log 2>&1
LS=130,200,256 RBN_Y_CACHE_GB=0 timeout 300 python tools/experiments/exp_long.py > gpurun_out/r2_x41_nocache.log 2>&1
N=128 LS=150,257 timeout 300 python tools/experiments/exp_long.py > gpurun_out/r2_x41_n128.log 2>&1
for f in gpurun_out/r2_x41_*.log; do echo "== $f"; cut -c1-330 $f | tail -12; done
