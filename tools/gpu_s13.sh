#!/bin/bash
mkdir -p gpurun_out
DIFFQC_B200_LIB=$PWD/variants/lib_loose.so SKIP_GENERIC=1 ENGINES=1 KG=4,5 timeout 300 python tools/quick_bench.py 2>&1 | tee gpurun_out/s13.txt
DIFFQC_B200_LIB=$PWD/variants/lib_loosetrace.so G=5 timeout 120 python tools/trace_phases.py 2>&1 | grep "items\|gap before\|busy\|warp \|late items\|warm items" | tee -a gpurun_out/s13.txt
