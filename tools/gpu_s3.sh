#!/bin/bash
# GPU session 3: parity tests, bench line, component-removal timings, phase trace, one full ncu capture
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/s3_smi.txt
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/s3_tests.log 2>&1
tail -3 gpurun_out/s3_tests.log
timeout 600 python bench.py > gpurun_out/s3_bench.json 2> gpurun_out/s3_bench.err
cat gpurun_out/s3_bench.json
echo "== baseline"; timeout 120 python tools/exp_grid.py 2>&1 | tee gpurun_out/s3_exp0.txt
for e in 1 2 4 3 7; do
  echo "== DQ_EXP=$e (1 no-FP64, 2 no-smem, 4 no-global)"
  DIFFQC_B200_LIB=$PWD/variants/lib_exp$e.so timeout 120 python tools/exp_grid.py 2>&1 | tee gpurun_out/s3_exp$e.txt
done
echo "== trace G=5"
DIFFQC_B200_LIB=$PWD/variants/lib_trace.so G=5 timeout 120 python tools/trace_phases.py 2>&1 | tee gpurun_out/s3_trace.txt
G=5 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fused_passes -s 6 -c 1 -f -o gpurun_out/s3_prof python tools/profile_case.py > gpurun_out/s3_ncu.log 2>&1
tail -3 gpurun_out/s3_ncu.log
