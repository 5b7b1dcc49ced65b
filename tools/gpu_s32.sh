#!/bin/bash
# GPU session 32: full ncu capture of the fused rotation pass (contiguous pass and one high-bit pass), n = 28 on one GPU
mkdir -p gpurun_out
N=28 STEPS=1 FUSED=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_slice_rx_tile -s 6 -c 2 -f -o gpurun_out/s32_rx_tile python tools/dist_state_run.py > gpurun_out/s32_ncu.log 2>&1
tail -2 gpurun_out/s32_ncu.log
