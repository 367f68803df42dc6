#!/bin/bash
# round 2, GPU call 33: experiment -- micro-batches dealt over two launch streams (tail filling across micro-batches)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
A="--steps 5 --warmup 3 --skip-e2e --skip-cpu --no-secondary"
timeout 300 python bench.py $A > gpurun_out/r2_b33_default.log 2>&1
RBN_WS_POOL=0 timeout 300 python bench.py $A --micro 64 > gpurun_out/r2_b33_m64_s1.log 2>&1
RBN_WS_POOL=0 timeout 300 python bench.py $A --micro 64 --streams 2 > gpurun_out/r2_b33_m64_s2.log 2>&1
RBN_WS_POOL=0 timeout 300 python bench.py $A --micro 32 --streams 4 > gpurun_out/r2_b33_m32_s4.log 2>&1
timeout 300 python bench.py $A > gpurun_out/r2_b33_default2.log 2>&1
python - <<'P'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2_b33_*.log')):
    ok = False
    for l in open(f):
        if l.startswith('{'):
            d = json.loads(l); ok = True
            print(f, round(d['value'], 1), round(d['ms_per_step'], 3), d['roofline'].get('kernels_ms_per_step'), d.get('clocks', {}).get('sm_mhz'), d.get('count_check'), d.get('fixture_check', {}).get('rel_err'))
    if not ok: print(f, 'NO LINE', open(f).read()[-400:])
P
