#!/bin/bash
# GPU session 30: Gray-code energy kernel: slice parity tests, launch list at n = 28
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_distributed_state.py -m gpu -x -q ) > gpurun_out/s30_tests.log 2>&1
head -3 gpurun_out/s30_tests.log
N=28 STEPS=2 FUSED=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_slice --csv --log-file gpurun_out/s30_fused.csv python tools/dist_state_run.py > gpurun_out/s30_fused.log 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/s30_fused.csv")) if len(r) > 10 and r[0].isdigit()]
for r in rows:
    print("%3s  %-60s %10.3f ms" % (r[0], r[4][:60], float(r[-1]) / 1e6))
PY
grep "energy\|norm2" gpurun_out/s30_fused.log
