#!/bin/bash
# round 2, GPU call 42: long-sequence tests (2-4 split windows) + the whole headline test file
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
LS=258,300,400,513 timeout 600 python tools/experiments/exp_long.py > gpurun_out/r2_x42_default.log 2>&1
cut -c1-330 gpurun_out/r2_x42_default.log | tail
timeout 900 python -m pytest tests/test_gpu_headline.py -q -m gpu > gpurun_out/r2_t42_headline.log 2>&1
tail -15 gpurun_out/r2_t42_headline.log | cut -c1-300
