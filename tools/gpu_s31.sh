#!/bin/bash
# GPU session 31 (2 GPUs): reference arm under torchrun after the OpenMP fix (expects cores = all host cores), short b200 arm
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/s31_bench_ref_n2.json 2> gpurun_out/s31_bench_ref_n2.err
cat gpurun_out/s31_bench_ref_n2.json | cut -c1-700
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/s31_bench_n2.json 2> gpurun_out/s31_bench_n2.err
cat gpurun_out/s31_bench_n2.json | cut -c1-300
