#!/bin/bash
# GPU session 36 (2 GPUs): stdout of the bench under torchrun must be exactly one JSON line
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/s36_bench_n2.json 2> gpurun_out/s36_bench_n2.err
wc -l gpurun_out/s36_bench_n2.json; cut -c1-120 gpurun_out/s36_bench_n2.json; grep -c "NCCL version" gpurun_out/s36_bench_n2.err
