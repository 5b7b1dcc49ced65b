#!/bin/bash
# GPU session 26: balanced fused rotation passes + Gray-code phase kernel: parity, step time at n = 30 (16 GiB slice), launch list at n = 28
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_distributed_state.py -m gpu -x -q ) > gpurun_out/s26_tests.log 2>&1
tail -4 gpurun_out/s26_tests.log
N=28 STEPS=2 FUSED=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_slice --csv --log-file gpurun_out/s26_fused.csv python tools/dist_state_run.py > gpurun_out/s26_fused.log 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/s26_fused.csv")) if len(r) > 10 and r[0].isdigit()]
for r in rows[-14:]:
    print(r[0], r[4][:44], r[-3], r[-2], r[-1])
PY
grep "energy_after\|norm2_after" gpurun_out/s26_fused.log
