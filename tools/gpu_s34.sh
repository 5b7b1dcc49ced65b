#!/bin/bash
# GPU session 34: full GPU suite + smoke at the final HEAD of round 1
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/s34_tests.log 2>&1
tail -5 gpurun_out/s34_tests.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1 | tee gpurun_out/s34_smoke.log
