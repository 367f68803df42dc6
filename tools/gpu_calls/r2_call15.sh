#!/bin/bash
# round 2, GPU call 15: GPU tests; K1 with 4 staging tiles (3 stores in flight) vs 2; K3 with 3 staging tiles per group
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest15.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest15.log
tail -4 gpurun_out/r2_pytest15.log
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 3 --warmup 2 --skip-cpu --skip-e2e --no-secondary > gpurun_out/r2_b15_$name.log 2>&1; }
run a_nbuf4 X=1
run b_nbuf2 RBN_TC_K1_NBUF=2
run c_nbuf4_gy3 RBN_TC_GY_NBUF=3
run d_nbuf4 X=1
run e_nbuf2 RBN_TC_K1_NBUF=2
run f_nbuf4_gy3 RBN_TC_GY_NBUF=3
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_b15_*.log')):
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line)
            print(f, 'value',round(d['value'],1),'ms',round(d['ms_per_step'],1),'frac',round(d['roofline']['step_frac'],4), d['clocks']['sm_mhz'], d['roofline']['kernels_ms_per_step'])
P
