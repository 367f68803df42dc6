#!/bin/bash
# round 2, GPU call 43: long sequences, tensor-core path vs CUDA-core path (time); then the whole GPU suite + smoke at HEAD
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python tools/experiments/exp_long_perf.py > gpurun_out/r2_x43_long_perf.log 2>&1
cat gpurun_out/r2_x43_long_perf.log | tail
( time timeout 1200 python -m pytest tests -q -m gpu -x ) > gpurun_out/r2_t43_all.log 2>&1
tail -8 gpurun_out/r2_t43_all.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
