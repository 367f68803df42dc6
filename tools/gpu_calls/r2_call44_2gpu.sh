#!/bin/bash
# round 2, GPU call 44 (2 GPUs): the driver's N = 2 command at HEAD (C3 + C5 secondary, NCCL), reference arm under torchrun
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_b44_2gpu.log 2>&1
echo "rc=$?"
python - <<'P'
import json
for l in open('gpurun_out/r2_b44_2gpu.log'):
    if l.startswith('{'):
        d = json.loads(l)
        print('C3', d['n_gpus'], round(d['value'], 1), round(d['ms_per_step'], 2), 'e2e', d['e2e']['value'], d['roofline'].get('step_frac'), d.get('clocks'))
        s = d.get('secondary') or {}
        print('C5', s.get('value'), s.get('ms_per_step'), (s.get('e2e') or {}).get('value'), (s.get('roofline') or {}).get('step_frac'), s.get('count_check'), s.get('error'))
P
tail -3 gpurun_out/r2_b44_2gpu.log | cut -c1-300
