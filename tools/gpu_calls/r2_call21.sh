#!/bin/bash
# round 2, GPU call 21: pair mode with Kpad-sized atoms (deeper ring on narrow diagonals): parity, C2 lines per ring depth
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_edges.py tests/test_gpu_headline.py -m gpu -q -x > gpurun_out/r2_pytest21.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest21.log
tail -4 gpurun_out/r2_pytest21.log
ARGS2="--workload c2 --steps 20 --warmup 5 --skip-e2e --skip-cpu --no-secondary"
for st in 0 2 3 4 6; do
  RBN_TC_K1_STAGES=$st timeout 300 python bench.py $ARGS2 > gpurun_out/r2_b21_c2_stages$st.log 2>&1
done
RBN_TC_K1_CTAS=1 timeout 300 python bench.py $ARGS2 > gpurun_out/r2_b21_c2_ctas1.log 2>&1
python - <<'P'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2_b21_c2_*.log')):
    for l in open(f):
        if l.startswith('{'):
            d = json.loads(l)
            print(f, round(d['value'], 1), round(d['ms_per_step'], 3), d['roofline'].get('kernels_ms_per_step'), d.get('gpu_launches'))
P
