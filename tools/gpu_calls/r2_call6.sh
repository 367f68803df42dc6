#!/bin/bash
# round 2, GPU call 6: GPU tests; accumulator-truncation compensation: kappa sweep on the C3 bench (fixture + count check)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest6.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest6.log
tail -5 gpurun_out/r2_pytest6.log
cp gpurun_out/parity_headline_c3.json gpurun_out/r2_parity6_c3.json; cp gpurun_out/parity_headline_c5.json gpurun_out/r2_parity6_c5.json
timeout 600 python bench.py --steps 4 --warmup 2 > gpurun_out/r2_b6_default.log 2>&1
echo "rc=$?" >> gpurun_out/r2_b6_default.log
for k in 0 3.0e-7 4.4e-7; do
  RBN_TC_KAPPA=$k timeout 300 python bench.py --steps 2 --warmup 1 --skip-cpu --skip-e2e > gpurun_out/r2_b6_kappa$k.log 2>&1
done
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_b6_*.log')):
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line)
            print(f, 'value',round(d['value'],1),'ms',round(d['ms_per_step'],1),'e2e',d['e2e']['value'],'frac',round(d['roofline']['step_frac'],4), d['clocks']['sm_mhz'], d['roofline']['kernels_ms_per_step'])
            print('    ', d.get('fixture_check'), d.get('count_check'), d.get('parity_check'))
P
