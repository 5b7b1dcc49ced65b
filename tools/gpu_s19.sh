#!/bin/bash
# GPU session 19 (2 GPUs): the sample-sharded bench line and the reference arm under torchrun, plus the 2-GPU tests that the
# single-GPU suite skips (distributed state over NCCL).
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/s19_bench_n2.json 2> gpurun_out/s19_bench_n2.err
cat gpurun_out/s19_bench_n2.json; tail -3 gpurun_out/s19_bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/s19_bench_ref_n2.json 2> gpurun_out/s19_bench_ref_n2.err
cat gpurun_out/s19_bench_ref_n2.json; tail -3 gpurun_out/s19_bench_ref_n2.err
( time timeout 600 python -m pytest tests/test_distributed_state.py tests/test_sharding.py -m gpu -x -q ) > gpurun_out/s19_tests_2gpu.log 2>&1
tail -5 gpurun_out/s19_tests_2gpu.log
