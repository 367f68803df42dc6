#!/bin/bash
# round 2, GPU call 14: Gaussian kernels, register budget A/B; default bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
RBN_GAUSS_OCC=1 timeout 600 python -m pytest tests/test_gpu_gauss.py -m gpu -q > gpurun_out/r2_pytest14.log 2>&1
echo "pytest(occ=1) rc=$?" >> gpurun_out/r2_pytest14.log; tail -3 gpurun_out/r2_pytest14.log
for o in 0 1 0 1; do
  RBN_GAUSS_OCC=$o timeout 300 python bench.py --workload c4 --steps 5 --warmup 3 --skip-cpu > gpurun_out/r2_b14_c4_occ$o.log 2>&1
  python - "$o" <<'P'
import json,sys
for line in open(f'gpurun_out/r2_b14_c4_occ{sys.argv[1]}.log'):
    if line.startswith('{'):
        d=json.loads(line); print('occ',sys.argv[1], round(d['value'],1), round(d['ms_per_step'],2), d['roofline']['kernels_ms_per_step'])
P
done
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_b14_default.log 2>&1
python - <<'P'
import json
for line in open('gpurun_out/r2_b14_default.log'):
    if line.startswith('{'):
        d=json.loads(line)
        print('value',round(d['value'],1),'ms',round(d['ms_per_step'],1),'e2e',d['e2e']['value'],'frac',round(d['roofline']['step_frac'],4), d['clocks'], d['roofline']['kernels_ms_per_step'])
        print(d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['traffic'], d['roofline'].get('algorithmic_per_launch'))
        print(d.get('fixture_check'), d.get('count_check')); print(d.get('parity_check'))
        for k,v in (d.get('other_workloads') or {}).items(): print(k, v['value'], v['ms_per_step'], v['kernels_ms_per_step'])
P
