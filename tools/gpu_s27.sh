#!/bin/bash
# GPU session 27 (8 GPUs): BASELINE configs[4], one n = 32 MaxCut state (64 GiB) split over 8 ranks, fused slice kernels
mkdir -p gpurun_out
N=32 STEPS=3 FUSED=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/dist_state_run.py > gpurun_out/s27_n32_w8.log 2>&1
tail -24 gpurun_out/s27_n32_w8.log
