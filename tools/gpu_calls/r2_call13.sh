#!/bin/bash
# round 2, GPU call 13: GPU tests (new D = 8 Gaussian kernels); C4 A/B; ncu full for the two k_tc_gemm2 instantiations
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest13.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest13.log
tail -15 gpurun_out/r2_pytest13.log
timeout 300 python bench.py --workload c4 --steps 5 --warmup 3 --skip-cpu > gpurun_out/r2_b13_c4_rows.log 2>&1; tail -c 700 gpurun_out/r2_b13_c4_rows.log
RBN_GAUSS_ROWS=0 timeout 300 python bench.py --workload c4 --steps 5 --warmup 3 --skip-cpu > gpurun_out/r2_b13_c4_old.log 2>&1; tail -c 700 gpurun_out/r2_b13_c4_old.log
ARGS="--batch 128 --micro 128 --steps 1 --warmup 1 --skip-e2e --skip-cpu --no-secondary"
full() {
    local name=$1 pat=$2 skip=$3; shift 3
    ncu --set full --clock-control none --kernel-name-base demangled -k "regex:$pat" -s $skip -c 1 -f -o gpurun_out/r02_full_$name \
        python bench.py "$@" > gpurun_out/r02_ncu_$name.log 2>&1
    echo "$name rc=$? $(tail -1 gpurun_out/r02_ncu_$name.log | cut -c1-200)"
    ncu -i gpurun_out/r02_full_$name.ncu-rep --page raw --csv > gpurun_out/r02_full_$name.csv 2>/dev/null
    rm -f gpurun_out/r02_full_$name.ncu-rep
}
python bench.py $ARGS > gpurun_out/r02_plain_c3b.log 2>&1 && {
full k2 "gemm2.*EpiRule" 172 $ARGS
full k4 "gemm2.*EpiCounts" 1 $ARGS
}
ARGS2="--workload c2 --steps 1 --warmup 1 --skip-e2e --skip-cpu --no-secondary"
python bench.py $ARGS2 > gpurun_out/r02_plain_c2b.log 2>&1 && full c2_k2 "gemm2.*EpiRule" 85 $ARGS2
ARGS4="--workload c4 --steps 1 --warmup 1 --skip-cpu --no-secondary"
python bench.py $ARGS4 > gpurun_out/r02_plain_c4b.log 2>&1 && {
full c4_fwd8 k_gauss_diag8 90 $ARGS4
full c4_bwd8 k_gauss_bwd_diag8 90 $ARGS4
}
ls -la gpurun_out/r02_full_k2.csv gpurun_out/r02_full_k4.csv gpurun_out/r02_full_c2_k2.csv gpurun_out/r02_full_c4_*8.csv
