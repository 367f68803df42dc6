#!/bin/bash
set -u
ARGS="--batch 128 --steps 1 --warmup 1 --skip-e2e --skip-cpu"
for spec in 299:k2 424:k4; do
    IFS=: read skip name <<< "$spec"
    ncu --set full --clock-control none --import-source on -k "regex:k_tc_gemm2" -s $skip -c 1 -f -o gpurun_out/prof5_$name \
        python bench.py $ARGS > gpurun_out/ncu5_$name.log 2>&1
    tail -1 gpurun_out/ncu5_$name.log
done
ARGS5="--workload c5 --batch 64 --steps 1 --warmup 1 --skip-e2e --skip-cpu"
ncu --set full --clock-control none --import-source on -k "regex:k_tc_gemm2" -s 299 -c 1 -f -o gpurun_out/prof5_k2n512 \
    python bench.py $ARGS5 > gpurun_out/ncu5_k2n512.log 2>&1
tail -1 gpurun_out/ncu5_k2n512.log
