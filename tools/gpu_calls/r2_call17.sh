#!/bin/bash
# round 2, GPU call 17: smoke(); the headline parity tests against the LIVE oracle over all N^3 entries; driver-length bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/r2_smoke.log
RBN_FULL_ORACLE=1 timeout 1500 python -m pytest tests/test_gpu_headline.py -m gpu -q -k "headline_shape" > gpurun_out/r2_full_oracle.log 2>&1; echo "full-oracle rc=$?"; tail -3 gpurun_out/r2_full_oracle.log
cp gpurun_out/parity_headline_c3.json gpurun_out/r2_parity_full_oracle_c3.json; cp gpurun_out/parity_headline_c5.json gpurun_out/r2_parity_full_oracle_c5.json
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_b17_driver_like.log 2>&1; echo "bench rc=$?"
python - <<'P'
import json
for line in open('gpurun_out/r2_b17_driver_like.log'):
    if line.startswith('{'):
        d=json.loads(line)
        print('value',round(d['value'],1),'ms',round(d['ms_per_step'],1),'e2e',d['e2e']['value'],'frac',round(d['roofline']['step_frac'],4), d['clocks'], d['roofline']['kernels_ms_per_step'])
        print(d.get('parity_check'))
P
