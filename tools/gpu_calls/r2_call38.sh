#!/bin/bash
# round 2, GPU call 38: C5 (N=512) with the two-stream schedules -- the HBM-bound split-sum / child-gradient kernels beside the tensor-bound GEMMs?
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
A="--workload c5 --steps 2 --warmup 1 --skip-cpu --skip-e2e --no-secondary"
for ov in 1 2 3; do
  RBN_TC_OVERLAP=$ov timeout 400 python bench.py $A > gpurun_out/r2_b38_c5_ov$ov.log 2>&1
done
RBN_TC_OVERLAP=1 RBN_TC_FWD_SMS0=60 RBN_TC_BWD_SMS0=60 timeout 400 python bench.py $A > gpurun_out/r2_b38_c5_ov1_sms60.log 2>&1
python - <<'P'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2_b38_*.log')):
    ok = False
    for l in open(f):
        if l.startswith('{'):
            d = json.loads(l); ok = True
            r = d['roofline']
            print(f, round(d['value'], 1), round(d['ms_per_step'], 3), r.get('step_frac'), r.get('kernels_ms_per_step'), r.get('split_sum_passes'), d.get('logZ_mean'), d.get('count_check'), d.get('clocks', {}).get('sm_mhz'))
    if not ok: print(f, 'NO LINE', open(f).read()[-300:])
P
