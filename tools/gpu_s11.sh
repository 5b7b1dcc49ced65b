#!/bin/bash
# GPU session 11: full ncu capture of the chained two-team kernel (one sample = one launch of 100 kets)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fused_passes -s 1 -c 1 -f -o gpurun_out/s11_prof python tools/profile_case.py > gpurun_out/s11_ncu.log 2>&1
tail -3 gpurun_out/s11_ncu.log
