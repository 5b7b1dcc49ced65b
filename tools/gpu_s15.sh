#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_ising_gpu.py -q -x ) > gpurun_out/s15_tests.log 2>&1
tail -4 gpurun_out/s15_tests.log
SKIP_GENERIC=1 ENGINES=1 KG=4,5,6 timeout 300 python tools/quick_bench.py 2>&1 | tee gpurun_out/s15_quick.txt
N=16 B=2 SKIP_GENERIC=1 ENGINES=1 KG=80 timeout 300 python tools/quick_bench.py 2>&1 | tee gpurun_out/s15_quick16.txt
DIFFQC_B200_LIB=$PWD/variants/lib_trace.so G=5 timeout 120 python tools/trace_phases.py 2>&1 | grep "items\|gap before\|busy\|warp \|late items\|warm items" | tee gpurun_out/s15_trace.txt
