#!/bin/bash
# GPU session: parity tests, bench line, component-removal timings, phase trace
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/s1_smi.txt
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/s1_tests.log 2>&1
tail -3 gpurun_out/s1_tests.log
timeout 600 python bench.py > gpurun_out/s1_bench.json 2> gpurun_out/s1_bench.err
cat gpurun_out/s1_bench.json
echo "== baseline"; timeout 120 python tools/exp_grid.py 2>&1 | tee gpurun_out/s1_exp0.txt
for e in 1 2 4 3 5 6 7; do
  echo "== DQ_EXP=$e (1 no-FP64, 2 no-smem, 4 no-global)"
  DIFFQC_B200_LIB=$PWD/variants/lib_exp$e.so timeout 120 python tools/exp_grid.py 2>&1 | tee gpurun_out/s1_exp$e.txt
done
echo "== trace G=5 gps=2"
DIFFQC_B200_LIB=$PWD/variants/lib_trace.so G=5 timeout 120 python tools/trace_phases.py 2>&1 | tee gpurun_out/s1_trace.txt
