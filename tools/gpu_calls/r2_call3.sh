#!/bin/bash
# round 2, GPU call 3: GPU tests; C3 bench after K2 short-unit mode + K3 uneven ranges; launch list (ncu) of one micro-batch
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest3.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest3.log
tail -5 gpurun_out/r2_pytest3.log
timeout 600 python bench.py --steps 3 --warmup 2 --skip-cpu > gpurun_out/r2_b3_default.log 2>&1
echo "rc=$?" >> gpurun_out/r2_b3_default.log
timeout 300 python bench.py --steps 3 --warmup 2 --micro 512 --skip-cpu --skip-e2e > gpurun_out/r2_b3_micro512.log 2>&1
timeout 300 python bench.py --steps 1 --warmup 1 --batch 128 --micro 128 --skip-cpu --skip-e2e > gpurun_out/r2_b3_plain128.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_b128.csv python bench.py --steps 1 --warmup 1 --batch 128 --micro 128 --skip-cpu --skip-e2e > gpurun_out/r2_ncu_ll.log 2>&1
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_b3_*.log')):
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line)
            print(f, 'value',round(d['value'],1),'ms',round(d['ms_per_step'],1),'e2e',d['e2e']['value'],'frac',round(d['roofline']['step_frac'],4), d['clocks']['sm_mhz'], d['roofline']['kernels_ms_per_step'])
P
