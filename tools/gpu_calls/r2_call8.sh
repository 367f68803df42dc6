#!/bin/bash
# round 2, GPU call 8: GPU tests; C3 bench (warp-per-span finalize), twice
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest8.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest8.log
tail -4 gpurun_out/r2_pytest8.log
for i in 1 2; do timeout 300 python bench.py --steps 3 --warmup 2 --skip-cpu --skip-e2e > gpurun_out/r2_b8_run$i.log 2>&1; done
timeout 300 python bench.py --steps 3 --warmup 2 --skip-cpu --skip-e2e --micro 512 > gpurun_out/r2_b8_micro512.log 2>&1
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_b8_*.log')):
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line)
            print(f, 'value',round(d['value'],1),'ms',round(d['ms_per_step'],1),'frac',round(d['roofline']['step_frac'],4), d['clocks']['sm_mhz'], d['roofline']['kernels_ms_per_step'])
P
