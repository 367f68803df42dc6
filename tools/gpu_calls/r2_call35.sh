#!/bin/bash
# round 2, GPU call 35: C2 (N=64, inside only) -- does an L2-resident Y ring (small launch groups, optional two-stream overlap) beat one launch per diagonal?
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
A="--workload c2 --steps 10 --warmup 3 --skip-cpu --skip-e2e --no-secondary"
timeout 120 python bench.py $A > gpurun_out/r2_b35_c2_default.log 2>&1
for ch in 4096 8192 16384 32768; do
  for ov in 0 1 2 3; do
    RBN_TC_OVERLAP=$ov timeout 120 python bench.py $A --chunk $ch > gpurun_out/r2_b35_c2_ch${ch}_ov${ov}.log 2>&1
  done
done
python - <<'P'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2_b35_*.log')):
    ok = False
    for l in open(f):
        if l.startswith('{'):
            d = json.loads(l); ok = True
            r = d['roofline']
            print(f, round(d['value'], 1), round(d['ms_per_step'], 3), r.get('kernels_ms_per_step'), d.get('logZ_mean'), d.get('gpu_launches'))
    if not ok: print(f, 'NO LINE', open(f).read()[-300:])
P
