#!/bin/bash
# round 2, GPU call 27: DRAM bytes of the N = 512 rule GEMM launches at HEAD (does the second parent-half tile hit L2?)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
ARGS="--workload c5 --batch 128 --micro 128 --steps 1 --warmup 1 --skip-e2e --skip-cpu --no-secondary"
python bench.py $ARGS > gpurun_out/r02_plain_c5_b128.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum --clock-control none -k regex:k_tc_gemm2 -s 40 -c 60 --csv --log-file gpurun_out/r02_launches_c5_k2.csv python bench.py $ARGS > gpurun_out/r02_ncu_c5_k2.log 2>&1
tail -3 gpurun_out/r02_plain_c5_b128.log | cut -c1-400
wc -l gpurun_out/r02_launches_c5_k2.csv
