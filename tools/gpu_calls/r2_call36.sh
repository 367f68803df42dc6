#!/bin/bash
# round 2, GPU call 36: programmatic dependent launch (RBN_PDL=1) -- parity tests, then C2 / C3 with and without it on the same box
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
RBN_PDL=1 timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_headline.py tests/test_gpu_parity.py tests/test_gpu_em.py -q -m gpu -x > gpurun_out/r2_t36_pdl.log 2>&1
tail -5 gpurun_out/r2_t36_pdl.log
A="--steps 10 --warmup 3 --skip-cpu --skip-e2e --no-secondary"
for p in 0 1 0 1; do
  RBN_PDL=$p timeout 120 python bench.py --workload c2 $A > gpurun_out/r2_b36_c2_pdl${p}_$RANDOM.log 2>&1
done
B="--steps 5 --warmup 3 --skip-cpu --skip-e2e --no-secondary"
RBN_PDL=0 timeout 300 python bench.py $B > gpurun_out/r2_b36_c3_pdl0_a.log 2>&1
RBN_PDL=1 timeout 300 python bench.py $B > gpurun_out/r2_b36_c3_pdl1_a.log 2>&1
RBN_PDL=0 timeout 300 python bench.py $B > gpurun_out/r2_b36_c3_pdl0_b.log 2>&1
RBN_PDL=1 timeout 300 python bench.py $B > gpurun_out/r2_b36_c3_pdl1_b.log 2>&1
python - <<'P'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2_b36_*.log')):
    ok = False
    for l in open(f):
        if l.startswith('{'):
            d = json.loads(l); ok = True
            r = d['roofline']
            print(f, round(d['value'], 1), round(d['ms_per_step'], 3), r.get('kernels_ms_per_step'), d.get('logZ_mean'), d.get('count_check'), d.get('fixture_check', {}).get('rel_err'), d.get('clocks', {}).get('sm_mhz'))
    if not ok: print(f, 'NO LINE', open(f).read()[-300:])
P
