#!/bin/bash
# Round-1 profiling pass (run under gpurun, one GPU).  Every profiled command first runs once WITHOUT ncu and must exit 0.
set -u
ARGS="--batch 128 --steps 1 --warmup 1 --skip-e2e --skip-cpu"
python bench.py $ARGS > gpurun_out/prof_plain_c3.log 2>&1 || { echo "plain run failed"; exit 1; }
# launch list of the timed step (the warm-up step's 1100-odd launches are skipped)
ncu --metrics gpu__time_duration.sum --clock-control none -s 1100 -c 1300 --csv --log-file gpurun_out/launches_r01c.csv \
    python bench.py $ARGS > gpurun_out/ncu_ll_c3.log 2>&1
# one full capture per kernel of the path, taken in the timed step (skip counts are per kernel name)
for spec in splitsum:300:k1 "gemm2.*EpiRule:170:k2" gy2:170:k3 "gemm2.*EpiCounts:170:k4" childgrad:170:k5; do
    IFS=: read pat skip name <<< "$spec"
    ncu --set full --clock-control none --import-source on -k "regex:$pat" -s $skip -c 1 -f -o gpurun_out/prof5_$name \
        python bench.py $ARGS > gpurun_out/ncu5_$name.log 2>&1
    tail -1 gpurun_out/ncu5_$name.log
done
# N = 512 (C5 shape, 64 sequences): rule GEMM and gY GEMM
ARGS5="--workload c5 --batch 64 --steps 1 --warmup 1 --skip-e2e --skip-cpu"
python bench.py $ARGS5 > gpurun_out/prof_plain_c5.log 2>&1 || { echo "plain c5 run failed"; exit 1; }
for spec in "gemm2.*EpiRule:170:k2n512" gy2:170:k3n512; do
    IFS=: read pat skip name <<< "$spec"
    ncu --set full --clock-control none --import-source on -k "regex:$pat" -s $skip -c 1 -f -o gpurun_out/prof5_$name \
        python bench.py $ARGS5 > gpurun_out/ncu5_$name.log 2>&1
    tail -1 gpurun_out/ncu5_$name.log
done
