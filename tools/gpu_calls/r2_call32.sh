#!/bin/bash
# round 2, GPU call 32: split-sum kernel with two epilogue groups, two staging tiles each (one group per accumulator half)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
RBN_TC_K1_GROUPS=2 timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_headline.py -m gpu -q -x > gpurun_out/r2_pytest32.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest32.log
tail -3 gpurun_out/r2_pytest32.log
for i in 1 2 3; do
  RBN_TC_K1_GROUPS=1 timeout 300 python bench.py --steps 6 --warmup 3 --skip-e2e --skip-cpu --no-secondary > gpurun_out/r2_b32_c3_g1_run$i.log 2>&1
  RBN_TC_K1_GROUPS=2 timeout 300 python bench.py --steps 6 --warmup 3 --skip-e2e --skip-cpu --no-secondary > gpurun_out/r2_b32_c3_g2_run$i.log 2>&1
done
python - <<'P'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2_b32_*.log')):
    for l in open(f):
        if l.startswith('{'):
            d = json.loads(l)
            print(f, round(d['value'], 1), round(d['ms_per_step'], 3), d['roofline'].get('kernels_ms_per_step'), d.get('clocks', {}).get('sm_mhz'))
P
