#!/bin/bash
# round 2, GPU call 18: micro-batch size sweep around 148 (= one span per SM and round in K1 / K5, 74 tiles for 74 CTA pairs)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for m in 128 148 128 148 171; do
  timeout 300 python bench.py --steps 3 --warmup 2 --micro $m --skip-cpu --skip-e2e --no-secondary > gpurun_out/r2_b18_micro$m.log 2>&1
  python - "$m" <<'P'
import json,sys
for line in open(f'gpurun_out/r2_b18_micro{sys.argv[1]}.log'):
    if line.startswith('{'):
        d=json.loads(line); print('micro',sys.argv[1],'value',round(d['value'],1),'ms',round(d['ms_per_step'],1),'frac',round(d['roofline']['step_frac'],4), d['clocks']['sm_mhz'], d['roofline']['kernels_ms_per_step'], d['roofline'].get('split_sum_passes'))
P
  tail -2 gpurun_out/r2_b18_micro$m.log | cut -c1-300 | grep -i "error\|Traceback" 
done
