#!/bin/bash
# round 2, GPU call 26: C2 with two sequence groups pipelined on two streams (tail filling) vs the default single stream
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
ARGS2="--workload c2 --steps 20 --warmup 5 --skip-e2e --skip-cpu --no-secondary"
run() { name=$1; shift; env "$@" timeout 300 python bench.py $ARGS2 > gpurun_out/r2_b26_c2_$name.log 2>&1; }
run default RBN_TC_OVERLAP=0
run ov3_g2 RBN_TC_OVERLAP=3 RBN_TC_GROUPS=2
run ov3_g4 RBN_TC_OVERLAP=3 RBN_TC_GROUPS=4
run ov1_g2_s84 RBN_TC_OVERLAP=1 RBN_TC_GROUPS=2 RBN_TC_FWD_SMS0=84
run ov1_g2_s74 RBN_TC_OVERLAP=1 RBN_TC_GROUPS=2 RBN_TC_FWD_SMS0=74
run ov2_g2 RBN_TC_OVERLAP=2 RBN_TC_GROUPS=2
python - <<'P'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2_b26_*.log')):
    ok = False
    for l in open(f):
        if l.startswith('{'):
            d = json.loads(l); ok = True
            print(f, round(d['value'], 1), round(d['ms_per_step'], 3), d['roofline'].get('kernels_ms_per_step'), d.get('parity_check'))
    if not ok: print(f, 'NO LINE', open(f).read()[-300:])
P
