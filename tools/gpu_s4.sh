#!/bin/bash
# GPU session 4: TMA tile loads -- parity first, then A/B timing against the LDGSTS build (variants/lib_v2.so)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/s4_smi.txt
( time timeout 900 python -m pytest tests/test_ising_gpu.py -q ) > gpurun_out/s4_tests.log 2>&1
tail -25 gpurun_out/s4_tests.log
echo "== TMA build"
ENGINES=1 KG=4,5,6 timeout 300 python tools/quick_bench.py 2>&1 | tee gpurun_out/s4_quick_tma.txt
echo "== LDGSTS build (v2)"
DIFFQC_B200_LIB=$PWD/variants/lib_v2.so SKIP_GENERIC=1 ENGINES=1 KG=4,5,6 timeout 300 python tools/quick_bench.py 2>&1 | tee gpurun_out/s4_quick_v2.txt
echo "== n=16"
N=16 B=2 SKIP_GENERIC=1 ENGINES=1 KG=40,80 timeout 300 python tools/quick_bench.py 2>&1 | tee gpurun_out/s4_quick16_tma.txt
DIFFQC_B200_LIB=$PWD/variants/lib_v2.so N=16 B=2 SKIP_GENERIC=1 ENGINES=1 KG=40,80 timeout 300 python tools/quick_bench.py 2>&1 | tee gpurun_out/s4_quick16_v2.txt
