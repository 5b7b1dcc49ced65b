#!/bin/bash
# GPU session 12: paced teams -- parity, timing, instruction-cache counters
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_ising_gpu.py -q -x ) > gpurun_out/s12_tests.log 2>&1
tail -4 gpurun_out/s12_tests.log
SKIP_GENERIC=1 ENGINES=1 KG=4,5,6 timeout 300 python tools/quick_bench.py 2>&1 | tee gpurun_out/s12_quick.txt
N=16 B=2 SKIP_GENERIC=1 ENGINES=1 KG=80 timeout 300 python tools/quick_bench.py 2>&1 | tee gpurun_out/s12_quick16.txt
G=5 timeout 600 ncu --metrics gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,sm__icc_request_hit_rate.pct,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum --clock-control none -k regex:k_fused_passes -s 1 -c 1 python tools/profile_case.py 2>&1 | grep -A8 "k_fused_passes" | tee gpurun_out/s12_icache.txt
