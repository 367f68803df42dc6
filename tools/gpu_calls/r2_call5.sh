#!/bin/bash
# round 2, GPU call 5: GPU tests; C3 bench with the deeper K5 ring and stream-K rule GEMM
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest5.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest5.log
tail -5 gpurun_out/r2_pytest5.log
timeout 600 python bench.py --steps 4 --warmup 2 --skip-cpu > gpurun_out/r2_b5_default.log 2>&1
echo "rc=$?" >> gpurun_out/r2_b5_default.log
timeout 300 python bench.py --steps 4 --warmup 2 --micro 512 --skip-cpu --skip-e2e > gpurun_out/r2_b5_micro512.log 2>&1
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_b5_*.log')):
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line)
            print(f, 'value',round(d['value'],1),'ms',round(d['ms_per_step'],1),'e2e',d['e2e']['value'],'frac',round(d['roofline']['step_frac'],4), d['clocks']['sm_mhz'], d['roofline']['kernels_ms_per_step'])
P
