#!/bin/bash
# round 2, GPU call 1: GPU test suite, then the C3 bench with the split-sum cache at several micro-batch sizes
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,memory.used --format=csv > gpurun_out/r2_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q --deselect "tests/test_gpu_headline.py::test_headline_shape_vs_oracle[headline_c5]" > gpurun_out/r2_pytest1.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest1.log
timeout 600 python bench.py --steps 3 --warmup 2 > gpurun_out/r2_b1_micro128.log 2>&1
echo "rc=$?" >> gpurun_out/r2_b1_micro128.log
timeout 300 python bench.py --steps 3 --warmup 2 --micro 512 --skip-cpu --skip-e2e > gpurun_out/r2_b1_micro512.log 2>&1
echo "rc=$?" >> gpurun_out/r2_b1_micro512.log
RBN_Y_CACHE_GB=0 timeout 300 python bench.py --steps 3 --warmup 2 --micro 512 --skip-cpu --skip-e2e > gpurun_out/r2_b1_micro512_nocache.log 2>&1
echo "rc=$?" >> gpurun_out/r2_b1_micro512_nocache.log
timeout 300 python bench.py --steps 3 --warmup 2 --micro 256 --skip-cpu --skip-e2e > gpurun_out/r2_b1_micro256.log 2>&1
echo "rc=$?" >> gpurun_out/r2_b1_micro256.log
RBN_Y_CACHE_GB=0 timeout 300 python bench.py --steps 3 --warmup 2 --micro 128 --skip-cpu --skip-e2e > gpurun_out/r2_b1_micro128_nocache.log 2>&1
echo "rc=$?" >> gpurun_out/r2_b1_micro128_nocache.log
tail -3 gpurun_out/r2_pytest1.log
for f in gpurun_out/r2_b1_*.log; do echo "== $f"; tail -c 600 $f; done
