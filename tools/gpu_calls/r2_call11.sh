#!/bin/bash
# round 2, GPU call 11: GPU tests; full default bench (with other_workloads); C2 alone
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest11.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest11.log
tail -4 gpurun_out/r2_pytest11.log
timeout 900 python bench.py --steps 4 --warmup 3 > gpurun_out/r2_b11_default.log 2>&1
echo "rc=$?" >> gpurun_out/r2_b11_default.log
tail -c 600 gpurun_out/r2_b11_default.log
timeout 300 python bench.py --impl reference --workload c1 --steps 3 --warmup 1 > gpurun_out/r2_b11_ref_c1.log 2>&1
tail -c 900 gpurun_out/r2_b11_ref_c1.log
python - <<'P'
import json
for line in open('gpurun_out/r2_b11_default.log'):
    if line.startswith('{'):
        d=json.loads(line)
        print('value',round(d['value'],1),'ms',round(d['ms_per_step'],1),'e2e',d['e2e']['value'],'frac',round(d['roofline']['step_frac'],4), d['clocks']['sm_mhz'], d['roofline']['kernels_ms_per_step'])
        print(d.get('fixture_check'), d.get('count_check')); print(d.get('parity_check'))
        for k,v in (d.get('other_workloads') or {}).items(): print(k, json.dumps(v))
P
