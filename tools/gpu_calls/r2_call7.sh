#!/bin/bash
# round 2, GPU call 7: GPU tests (count-GEMM operand under a data-derived scale); A/B in one box: stream-K on/off, K5 ring 4 vs 8
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest7.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest7.log
tail -5 gpurun_out/r2_pytest7.log
cp gpurun_out/parity_headline_c3.json gpurun_out/r2_parity7_c3.json; cp gpurun_out/parity_headline_c5.json gpurun_out/r2_parity7_c5.json
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 3 --warmup 2 --skip-cpu --skip-e2e > gpurun_out/r2_b7_$name.log 2>&1; }
run a_default X=1
run b_nostreamk RBN_TC_STREAMK=0
run c_k5stages4 RBN_TC_K5_STAGES=4
run d_default2 X=1
run e_nostreamk2 RBN_TC_STREAMK=0
run f_k5stages4b RBN_TC_K5_STAGES=4
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_b7_*.log')):
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line)
            print(f, 'value',round(d['value'],1),'ms',round(d['ms_per_step'],1),'frac',round(d['roofline']['step_frac'],4), d['clocks']['sm_mhz'], d['roofline']['kernels_ms_per_step'])
            print('    ', d.get('fixture_check'), d.get('count_check'))
P
