#!/bin/bash
# round 2, GPU call 34: Gaussian path with float32 charts (RBN_GAUSS_F32) -- tests, then C4 in both chart dtypes and register budgets
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_gauss.py -q -m gpu -x > gpurun_out/r2_t34_gauss.log 2>&1
tail -15 gpurun_out/r2_t34_gauss.log
A="--steps 10 --warmup 3 --skip-cpu"
timeout 200 python bench.py --workload c4 $A > gpurun_out/r2_b34_c4_f64.log 2>&1
timeout 200 python bench.py --workload c4f32 $A > gpurun_out/r2_b34_c4_f32_occ1.log 2>&1
RBN_GAUSS_OCC=2 timeout 200 python bench.py --workload c4f32 $A > gpurun_out/r2_b34_c4_f32_occ2.log 2>&1
RBN_GAUSS_OCC=0 timeout 200 python bench.py --workload c4f32 $A > gpurun_out/r2_b34_c4_f32_occ0.log 2>&1
python - <<'P'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2_b34_*.log')):
    ok = False
    for l in open(f):
        if l.startswith('{'):
            d = json.loads(l); ok = True
            r = d['roofline']
            print(f, round(d['value'], 1), round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value'], 1), 'frac', r.get('frac'), r.get('kernels_ms_per_step'), d.get('logZ_mean'), d.get('clocks', {}).get('sm_mhz'))
    if not ok: print(f, 'NO LINE', open(f).read()[-600:])
P
