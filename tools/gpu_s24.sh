#!/bin/bash
# GPU session 24: fused multi-qubit rotation pass for distributed-state slices: parity, then per-step time at the slice size of
# BASELINE configs[4] (n = 30 on one GPU = 16 GiB slice, twice the per-rank size of n = 32 on 8 GPUs), per-term kernels vs fused passes
mkdir -p gpurun_out
N=30 STEPS=3 FUSED=0 timeout 600 python tools/dist_state_run.py > gpurun_out/s24_n30_perterm.log 2>&1; tail -22 gpurun_out/s24_n30_perterm.log
N=30 STEPS=3 FUSED=1 timeout 600 python tools/dist_state_run.py > gpurun_out/s24_n30_fused.log 2>&1; tail -22 gpurun_out/s24_n30_fused.log
