#!/bin/bash
# round 2, GPU call 37: C5 (N=512, 1024 sequences per GPU) -- micro-batches small enough that ALL split-sums stay cached (no K1 in the outside pass)?
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
A="--workload c5 --steps 2 --warmup 1 --skip-cpu --skip-e2e --no-secondary"
for m in 512 128 64 32; do
  timeout 400 python bench.py $A --micro $m > gpurun_out/r2_b37_c5_m$m.log 2>&1
done
python - <<'P'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2_b37_*.log')):
    ok = False
    for l in open(f):
        if l.startswith('{'):
            d = json.loads(l); ok = True
            r = d['roofline']
            print(f, round(d['value'], 1), round(d['ms_per_step'], 3), r.get('step_frac'), r.get('kernels_ms_per_step'), r.get('split_sum_passes'), d.get('logZ_mean'), d.get('count_check'), d.get('clocks', {}).get('sm_mhz'))
    if not ok: print(f, 'NO LINE', open(f).read()[-300:])
P
