#!/bin/bash
# GPU session 23: resident dense kernel after the ILP / bank / occupancy changes
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_dense_gpu.py -m gpu -x -q ) > gpurun_out/s23_tests.log 2>&1
tail -5 gpurun_out/s23_tests.log
timeout 300 python tools/dense_small_case.py 2>&1 | tee gpurun_out/s23_dense_small.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_small -s 3 -c 1 -f -o gpurun_out/s23_small_prof python tools/dense_small_case.py > gpurun_out/s23_ncu.log 2>&1
tail -2 gpurun_out/s23_ncu.log
