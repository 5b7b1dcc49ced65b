#!/bin/bash
mkdir -p gpurun_out
echo "== evict_last stores"
DIFFQC_B200_LIB=$PWD/variants/lib_evl.so SKIP_GENERIC=1 ENGINES=1 KG=4,5,6 timeout 300 python tools/quick_bench.py 2>&1 | tee gpurun_out/s17.txt
DIFFQC_B200_LIB=$PWD/variants/lib_evl.so G=5 timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_fused_passes -s 1 -c 1 python tools/profile_case.py 2>&1 | grep "dram__\|gpu__time" | tee -a gpurun_out/s17.txt
echo "== default"
SKIP_GENERIC=1 ENGINES=1 KG=5 timeout 300 python tools/quick_bench.py 2>&1 | tee -a gpurun_out/s17.txt
