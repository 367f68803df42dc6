#!/bin/bash
# round 2, GPU call 22: split-sum kernel templated on pair mode + division-free span iterator: parity, C2 and C3 lines
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_edges.py tests/test_gpu_headline.py tests/test_gpu_em.py -m gpu -q -x > gpurun_out/r2_pytest22.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest22.log
tail -4 gpurun_out/r2_pytest22.log
ARGS2="--workload c2 --steps 20 --warmup 5 --skip-e2e --skip-cpu --no-secondary"
for i in 1 2; do
  timeout 300 python bench.py $ARGS2 > gpurun_out/r2_b22_c2_run$i.log 2>&1
done
timeout 600 python bench.py --steps 6 --warmup 3 --skip-e2e --skip-cpu --no-secondary > gpurun_out/r2_b22_c3.log 2>&1
python - <<'P'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2_b22_*.log')):
    for l in open(f):
        if l.startswith('{'):
            d = json.loads(l)
            print(f, round(d['value'], 1), round(d['ms_per_step'], 3), d['roofline'].get('kernels_ms_per_step'), d.get('clocks', {}).get('sm_mhz'))
P
