#!/bin/bash
# Round-2 profiling pass (run under gpurun, one GPU).  Every profiled command first runs once WITHOUT ncu and must exit 0.
# Outputs (gpurun_out/): r02_launches_c3_b128.csv (every launch of one 128-sequence inside+outside pass: duration + DRAM
# bytes), r02_full_<kernel>.csv (ncu --set full, raw page, one launch per kernel), plain-run logs.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
set -u
ARGS="--batch 128 --micro 128 --steps 1 --warmup 1 --skip-e2e --skip-cpu --no-secondary"
python bench.py $ARGS > gpurun_out/r02_plain_c3.log 2>&1 || { echo "plain c3 run failed"; tail -5 gpurun_out/r02_plain_c3.log; exit 1; }
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/r02_launches_c3_b128.csv python bench.py $ARGS > gpurun_out/r02_ncu_ll.log 2>&1
echo "launch list rc=$?"
full() {   # full <name> <regex> <skip> <bench args...>
    local name=$1 pat=$2 skip=$3; shift 3
    ncu --set full --clock-control none --import-source on -k "regex:$pat" -s $skip -c 1 -f -o gpurun_out/r02_full_$name \
        python bench.py "$@" > gpurun_out/r02_ncu_$name.log 2>&1
    echo "$name rc=$? $(tail -1 gpurun_out/r02_ncu_$name.log)"
    ncu -i gpurun_out/r02_full_$name.ncu-rep --page raw --csv > gpurun_out/r02_full_$name.csv 2>/dev/null
    ncu -i gpurun_out/r02_full_$name.ncu-rep --page details --csv > gpurun_out/r02_details_$name.csv 2>/dev/null
    case $name in k1|k5) ;; *) rm -f gpurun_out/r02_full_$name.ncu-rep ;; esac     # keep two reports with source, drop the rest (64 MiB limit)
}
# one launch per kernel in the timed pass (the warm-up pass has 127 launches of each per-diagonal kernel): inside w = 47, outside w = 47
full k1 splitsum 172 $ARGS
full k2 "gemm2.*EpiRule" 172 $ARGS
full k3 gy2 208 $ARGS
full k4 "gemm2.*EpiCounts" 1 $ARGS
full k5 childgrad 208 $ARGS
full prep k_prep_g 208 $ARGS
full finalize k_rule_finalize 172 $ARGS
# C2 (N = 64, L = 64, inside): split-sum and rule GEMM at w = 24 of the timed pass
ARGS2="--workload c2 --steps 1 --warmup 1 --skip-e2e --skip-cpu --no-secondary"
python bench.py $ARGS2 > gpurun_out/r02_plain_c2.log 2>&1 || { echo "plain c2 run failed"; exit 1; }
full c2_k1 splitsum 85 $ARGS2
full c2_k2 "gemm2.*EpiRule" 85 $ARGS2
# C1 (CUDA-core path, one cooperative launch per pass) and C4 (Gaussian)
ARGS1="--workload c1 --steps 1 --warmup 1 --skip-e2e --skip-cpu --no-secondary"
python bench.py $ARGS1 > gpurun_out/r02_plain_c1.log 2>&1 || { echo "plain c1 run failed"; exit 1; }
full c1_inside k_inside_all 1 $ARGS1
full c1_outside k_outside_all 1 $ARGS1
ARGS4="--workload c4 --steps 1 --warmup 1 --skip-cpu --no-secondary"
python bench.py $ARGS4 > gpurun_out/r02_plain_c4.log 2>&1 || { echo "plain c4 run failed"; exit 1; }
full c4_fwd k_gauss_diag 90 $ARGS4
full c4_bwd k_gauss_bwd_diag 90 $ARGS4
ls -la gpurun_out/r02_* | head -60
