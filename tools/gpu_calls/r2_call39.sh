#!/bin/bash
# round 2, GPU call 39: spans wider than 129 tokens on the tensor-core path (windows of 128 splits + k_y_combine)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_headline.py -q -m gpu -x -k "longer_than_129 or length_cap or split_sum_cache" > gpurun_out/r2_t39_long.log 2>&1
tail -30 gpurun_out/r2_t39_long.log
