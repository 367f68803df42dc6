#!/bin/bash
# round 2, GPU call 25: pair-mode fix-up with a two-deep, branch-free exponent prefetch
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_edges.py -m gpu -q -x > gpurun_out/r2_pytest25.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest25.log
tail -3 gpurun_out/r2_pytest25.log
ARGS2="--workload c2 --steps 20 --warmup 5 --skip-e2e --skip-cpu --no-secondary"
for st in 0 0 2; do
  RBN_TC_K1_STAGES=$st timeout 300 python bench.py $ARGS2 > gpurun_out/r2_b25_c2_stages$st.$RANDOM.log 2>&1
done
python - <<'P'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2_b25_*.log')):
    for l in open(f):
        if l.startswith('{'):
            d = json.loads(l)
            print(f, round(d['value'], 1), round(d['ms_per_step'], 3), d['roofline'].get('kernels_ms_per_step'), d.get('clocks', {}).get('sm_mhz'))
P
