#!/bin/bash
# round 2, GPU call 23: source-level ncu of the pair-mode split-sum kernel (N = 64), one mid-width diagonal
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
ARGS2="--workload c2 --steps 1 --warmup 1 --skip-e2e --skip-cpu --no-secondary"
python bench.py $ARGS2 > gpurun_out/r02_plain_c2d.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:splitsum -s 85 -c 1 -f -o gpurun_out/r02_src_c2_k1_pair python bench.py $ARGS2 > gpurun_out/r02_ncu_src_c2_k1_pair.log 2>&1
ncu -i gpurun_out/r02_src_c2_k1_pair.ncu-rep --page source --csv > gpurun_out/r02_src_c2_k1_pair_source.csv 2>/dev/null
ncu -i gpurun_out/r02_src_c2_k1_pair.ncu-rep --page raw --csv > gpurun_out/r02_full_c2_k1_pair.csv 2>/dev/null
ls -la gpurun_out/r02_src_c2_k1_pair*
