#!/bin/bash
# round 2, GPU call 46: the GPU suite at the final test layout (long-sequence cases against the CUDA-core path)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -q -m gpu -x --durations=8 ) > gpurun_out/r2_t46_all.log 2>&1
tail -22 gpurun_out/r2_t46_all.log | cut -c1-200
