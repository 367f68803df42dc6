#!/bin/bash
# round 2, GPU call 16 (8 GPUs): the scaling line with its C5 `secondary` over NCCL, and the reference arm under torchrun
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${1:-8}
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_b16_${N}gpu.log 2>&1
echo "rc=$?" >> gpurun_out/r2_b16_${N}gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus $N --steps 1 --warmup 1 > gpurun_out/r2_b16_${N}gpu_ref.log 2>&1
echo "rc=$?" >> gpurun_out/r2_b16_${N}gpu_ref.log
python - "$N" <<'P'
import json,sys
n=sys.argv[1]
for f in (f'gpurun_out/r2_b16_{n}gpu.log', f'gpurun_out/r2_b16_{n}gpu_ref.log'):
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line)
            print(f, 'value',round(d['value'],3),'ms',round(d['ms_per_step'],1),'e2e',d['e2e']['value'], d.get('config'))
            if 'secondary' in d:
                s=d['secondary']; print('  secondary', round(s['value'],1), round(s['ms_per_step'],1), s['e2e']['value'], round(s['roofline']['step_frac'],4), s['config'])
    print(open(f).read()[-200:])
P
