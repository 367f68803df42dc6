#!/bin/bash
# round 2, GPU call 45: the driver's N = 1 commands at HEAD: default bench (20 steps, 5 warm-ups) and the reference arm
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_b45_driver_like.log 2>&1
echo "rc=$?"
timeout 300 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/r2_b45_reference_arm.log 2>&1
echo "ref rc=$?"
python - <<'P'
import json
for f in ('gpurun_out/r2_b45_driver_like.log', 'gpurun_out/r2_b45_reference_arm.log'):
    for l in open(f):
        if l.startswith('{'):
            d = json.loads(l)
            r = d.get('roofline') or {}
            print(f, round(d['value'], 3), round(d['ms_per_step'], 3), 'e2e', d.get('e2e', {}).get('value'), 'step_frac', r.get('step_frac'), r.get('bound'), r.get('frac'), r.get('kernels_ms_per_step'), d.get('clocks'), d.get('parity_check'), {k: (v.get('value'), v.get('ms_per_step'), v.get('error')) for k, v in (d.get('other_workloads') or {}).items()})
P
