#!/bin/bash
# round 2, GPU call 19: GPU tests (new: fp64 lognormalize, two terminal variables); source-level ncu of the N = 64 split-sum
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest19.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest19.log
tail -4 gpurun_out/r2_pytest19.log
ARGS2="--workload c2 --steps 1 --warmup 1 --skip-e2e --skip-cpu --no-secondary"
python bench.py $ARGS2 > gpurun_out/r02_plain_c2c.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:splitsum -s 85 -c 1 -f -o gpurun_out/r02_src_c2_k1 python bench.py $ARGS2 > gpurun_out/r02_ncu_src_c2_k1.log 2>&1
ncu -i gpurun_out/r02_src_c2_k1.ncu-rep --page source --csv > gpurun_out/r02_src_c2_k1_source.csv 2>/dev/null
ls -la gpurun_out/r02_src_c2_k1*
