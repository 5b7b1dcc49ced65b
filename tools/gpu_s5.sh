#!/bin/bash
# GPU session 5: bench line, phase trace and one full ncu capture of the TMA-fed pass kernel
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/s5_bench.json 2> gpurun_out/s5_bench.err
cat gpurun_out/s5_bench.json
echo "== trace G=5"
DIFFQC_B200_LIB=$PWD/variants/lib_trace.so G=5 timeout 120 python tools/trace_phases.py 2>&1 | tee gpurun_out/s5_trace.txt
G=5 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fused_passes -s 6 -c 1 -f -o gpurun_out/s5_prof python tools/profile_case.py > gpurun_out/s5_ncu.log 2>&1
tail -3 gpurun_out/s5_ncu.log
