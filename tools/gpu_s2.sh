#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_ising_gpu.py -m gpu -x -q ) > gpurun_out/s2_tests.log 2>&1
tail -3 gpurun_out/s2_tests.log
echo "== P1"; timeout 120 python tools/exp_grid.py 2>&1 | tee gpurun_out/s2_exp_p1.txt
B=4 SKIP_GENERIC=1 ENGINES=1 GROUPS=4,5,6 timeout 200 python tools/quick_bench.py 2>&1 | tee gpurun_out/s2_quick.txt
G=5 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fused_passes -s 6 -c 1 -f -o gpurun_out/s2_prof python tools/profile_case.py > gpurun_out/s2_ncu.log 2>&1
tail -3 gpurun_out/s2_ncu.log
