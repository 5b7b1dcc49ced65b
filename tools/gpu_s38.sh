#!/bin/bash
# GPU session 38: full ncu capture of the Gray-code phase pass and of one high-bit rotation pass, n = 28 on one GPU
mkdir -p gpurun_out
N=28 STEPS=1 FUSED=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_slice_phase_gray|k_slice_rx_tile" -s 7 -c 3 -f -o gpurun_out/s38_slice python tools/dist_state_run.py > gpurun_out/s38_ncu.log 2>&1
tail -2 gpurun_out/s38_ncu.log
