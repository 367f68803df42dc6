#!/bin/bash
# round 2, GPU call 2: full GPU test suite; C3 bench with deferred count GEMM; L2-hint sweep; micro-batch sweep
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest2.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest2.log
tail -5 gpurun_out/r2_pytest2.log
timeout 600 python bench.py --steps 3 --warmup 2 > gpurun_out/r2_b2_default.log 2>&1
echo "rc=$?" >> gpurun_out/r2_b2_default.log
for h in 1 2 4 8 15; do
  RBN_TC_HINTS=$h timeout 300 python bench.py --steps 3 --warmup 2 --skip-cpu --skip-e2e > gpurun_out/r2_b2_hints$h.log 2>&1
  echo "rc=$?" >> gpurun_out/r2_b2_hints$h.log
done
for m in 160 192 256 512; do
  timeout 300 python bench.py --steps 3 --warmup 2 --micro $m --skip-cpu --skip-e2e > gpurun_out/r2_b2_micro$m.log 2>&1
  echo "rc=$?" >> gpurun_out/r2_b2_micro$m.log
done
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_b2_*.log')):
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line)
            print(f, 'value',round(d['value'],1),'ms',round(d['ms_per_step'],1),'e2e',d['e2e']['value'],'frac',round(d['roofline']['step_frac'],4), d['clocks']['sm_mhz'], d['roofline']['kernels_ms_per_step'])
P
