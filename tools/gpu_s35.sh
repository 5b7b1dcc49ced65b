#!/bin/bash
# GPU session 35: n = 30 (16 GiB slice, one GPU) step time at the final HEAD, per-term kernels beside it
mkdir -p gpurun_out
N=30 STEPS=3 FUSED=1 timeout 300 python tools/dist_state_run.py > gpurun_out/s35_n30_fused.log 2>&1; grep "seconds_per_step\|launches\|norm2_after\|energy_after" gpurun_out/s35_n30_fused.log
