#!/bin/bash
# GPU session 7: two teams x three tile buffers, TMA stores -- parity, A/B timing vs the one-buffer TMA build, trace
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_ising_gpu.py -q -x ) > gpurun_out/s7_tests.log 2>&1
tail -12 gpurun_out/s7_tests.log
SKIP_GENERIC=1 ENGINES=1 KG=4,5,6,8 timeout 300 python tools/quick_bench.py 2>&1 | tee gpurun_out/s7_quick.txt
DIFFQC_B200_LIB=$PWD/variants/lib_tma1.so SKIP_GENERIC=1 ENGINES=1 KG=5 timeout 300 python tools/quick_bench.py 2>&1 | tee gpurun_out/s7_quick_tma1.txt
N=16 B=2 SKIP_GENERIC=1 ENGINES=1 KG=80 timeout 300 python tools/quick_bench.py 2>&1 | tee gpurun_out/s7_quick16.txt
DIFFQC_B200_LIB=$PWD/variants/lib_trace.so G=5 timeout 120 python tools/trace_phases.py 2>&1 | tee gpurun_out/s7_trace.txt
