#!/bin/bash
# round 2, GPU call 28: source-level ncu of the child-gradient kernel on a NARROW diagonal (w = 8; per-span time there is a
# flat 5.85 us per SM against a 3.7 us HBM floor)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
ARGS="--batch 128 --micro 128 --steps 1 --warmup 1 --skip-e2e --skip-cpu --no-secondary"
timeout 300 python bench.py $ARGS > gpurun_out/r02_plain_c3b.log 2>&1 || { echo "plain run failed"; exit 1; }
timeout 500 ncu --set full --clock-control none --import-source on -k regex:childgrad -s 247 -c 1 -f -o gpurun_out/r02_src_k5_w8 python bench.py $ARGS > gpurun_out/r02_ncu_src_k5_w8.log 2>&1
echo "ncu rc=$?"
timeout 120 ncu -i gpurun_out/r02_src_k5_w8.ncu-rep --page source --csv > gpurun_out/r02_src_k5_w8_source.csv 2>/dev/null
timeout 120 ncu -i gpurun_out/r02_src_k5_w8.ncu-rep --page raw --csv > gpurun_out/r02_full_k5_w8.csv 2>/dev/null
ls -la gpurun_out/r02_src_k5_w8* gpurun_out/r02_full_k5_w8.csv
