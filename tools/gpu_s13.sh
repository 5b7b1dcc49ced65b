#!/bin/bash
mkdir -p gpurun_out
for lib in early; do
  echo "== $lib"
  DIFFQC_B200_LIB=$PWD/variants/lib_$lib.so SKIP_GENERIC=1 ENGINES=1 KG=4,5 timeout 300 python tools/quick_bench.py 2>&1 | tee -a gpurun_out/s13.txt
  DIFFQC_B200_LIB=$PWD/variants/lib_$lib.so G=5 timeout 600 ncu --metrics gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,sm__icc_request_hit_rate.pct,gpu__time_duration.sum --clock-control none -k regex:k_fused_passes -s 1 -c 1 python tools/profile_case.py 2>&1 | grep "gcc__\|gpu__time\|icc" | tee -a gpurun_out/s13.txt
done
