#!/bin/bash
# round 2, GPU call 30: same, shared-memory address masked to 18 bits for cluster launches; full GPU suite
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest30.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest30.log
tail -3 gpurun_out/r2_pytest30.log
for i in 1 2; do
  timeout 600 python bench.py --steps 6 --warmup 3 --skip-e2e --skip-cpu --no-secondary > gpurun_out/r2_b30_c3_run$i.log 2>&1
done
python - <<'P'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2_b30_*.log')):
    for l in open(f):
        if l.startswith('{'):
            d = json.loads(l)
            print(f, round(d['value'], 1), round(d['ms_per_step'], 3), d['roofline'].get('kernels_ms_per_step'), d.get('clocks', {}).get('sm_mhz'))
P
