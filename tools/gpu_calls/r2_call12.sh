#!/bin/bash
# round 2, GPU call 12: GPU tests; K1 with two epilogue groups vs one, alternating in the same box
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest12.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest12.log
tail -4 gpurun_out/r2_pytest12.log
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 3 --warmup 2 --skip-cpu --skip-e2e --no-secondary > gpurun_out/r2_b12_$name.log 2>&1; }
run a_two X=1
run b_one RBN_TC_K1_GROUPS=1
run c_two X=1
run d_one RBN_TC_K1_GROUPS=1
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_b12_*.log')):
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line)
            print(f, 'value',round(d['value'],1),'ms',round(d['ms_per_step'],1),'frac',round(d['roofline']['step_frac'],4), d['clocks']['sm_mhz'], d['roofline']['kernels_ms_per_step'])
P
