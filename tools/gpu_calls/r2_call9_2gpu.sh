#!/bin/bash
# round 2, GPU call 9 (2 GPUs): the N > 1 bench line with its C5 `secondary`, over NCCL
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/r2_b9_2gpu.log 2>&1
echo "rc=$?" >> gpurun_out/r2_b9_2gpu.log
tail -c 3000 gpurun_out/r2_b9_2gpu.log
