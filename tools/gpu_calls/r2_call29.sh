#!/bin/bash
# round 2, GPU call 29: child-gradient kernel with template-based MMA descriptors (fewer instructions in the issuing thread)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_headline.py -m gpu -q -x > gpurun_out/r2_pytest29.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest29.log
tail -3 gpurun_out/r2_pytest29.log
for i in 1 2; do
  timeout 600 python bench.py --steps 6 --warmup 3 --skip-e2e --skip-cpu --no-secondary > gpurun_out/r2_b29_c3_run$i.log 2>&1
done
python - <<'P'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2_b29_*.log')):
    for l in open(f):
        if l.startswith('{'):
            d = json.loads(l)
            print(f, round(d['value'], 1), round(d['ms_per_step'], 3), d['roofline'].get('kernels_ms_per_step'), d.get('clocks', {}).get('sm_mhz'))
P
