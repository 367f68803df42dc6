#!/bin/bash
# round 2, GPU call 31: the driver's commands at HEAD: smoke(), default bench (20 steps, 5 warm-ups), C2 line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke OK')" > gpurun_out/r2_smoke31.log 2>&1
tail -2 gpurun_out/r2_smoke31.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_b31_driver_like.log 2>&1
timeout 300 python bench.py --workload c2 --steps 20 --warmup 5 --no-secondary > gpurun_out/r2_b31_c2.log 2>&1
python - <<'P'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2_b31_*.log')):
    for l in open(f):
        if l.startswith('{'):
            d = json.loads(l)
            r = d['roofline']
            print(f, round(d['value'], 1), round(d['ms_per_step'], 3), 'e2e', d.get('e2e', {}).get('value'), 'step_frac', r.get('step_frac'), r.get('bound'), r.get('frac'), r.get('kernels_ms_per_step'), d.get('clocks'), d.get('parity_check'), {k: (v.get('value'), v.get('ms_per_step')) for k, v in (d.get('other_workloads') or {}).items()})
P
