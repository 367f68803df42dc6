#!/bin/bash
# round 2, GPU call 20: pair mode of the N = 64 split-sum (two spans per MMA): parity tests, then C2 bench lines
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_edges.py tests/test_gpu_headline.py tests/test_gpu_em.py -m gpu -q -x > gpurun_out/r2_pytest20.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest20.log
tail -6 gpurun_out/r2_pytest20.log
ARGS2="--workload c2 --steps 20 --warmup 5 --skip-e2e --skip-cpu --no-secondary"
for i in 1 2; do
  timeout 300 python bench.py $ARGS2 > gpurun_out/r2_b20_c2_run$i.log 2>&1
done
for ch in 4096 8192 16384; do
  timeout 300 python bench.py $ARGS2 --chunk $ch > gpurun_out/r2_b20_c2_chunk$ch.log 2>&1
done
python - <<'P'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2_b20_c2_*.log')):
    for l in open(f):
        if l.startswith('{'):
            d = json.loads(l)
            print(f, round(d['value'], 1), round(d['ms_per_step'], 3), d['roofline'].get('kernels_ms_per_step'), d.get('gpu_launches'))
P
