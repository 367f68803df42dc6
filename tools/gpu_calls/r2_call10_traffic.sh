#!/bin/bash
# round 2, GPU call 10: DRAM bytes of EVERY launch of one 128-sequence inside+outside pass (C3 shapes) -> per-step traffic
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
ARGS="--batch 128 --micro 128 --steps 1 --warmup 1 --skip-e2e --skip-cpu"
timeout 300 python bench.py $ARGS > gpurun_out/r2_traffic_plain.log 2>&1 &&
timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/r2_traffic_b128.csv python bench.py $ARGS > gpurun_out/r2_traffic_ncu.log 2>&1
echo "rc=$?"
tail -2 gpurun_out/r2_traffic_ncu.log
