#!/bin/bash
# GPU session 25: per-launch durations of the slice kernels (n = 28 on one GPU, 4 GiB slice), fused vs per-term
mkdir -p gpurun_out
N=28 STEPS=2 FUSED=1 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_slice --csv --log-file gpurun_out/s25_fused.csv python tools/dist_state_run.py > gpurun_out/s25_fused.log 2>&1
N=28 STEPS=1 FUSED=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_slice --csv --log-file gpurun_out/s25_perterm.csv python tools/dist_state_run.py > gpurun_out/s25_perterm.log 2>&1
python - <<'PY'
import csv
for f in ("gpurun_out/s25_fused.csv", "gpurun_out/s25_perterm.csv"):
    rows = [r for r in csv.reader(open(f)) if len(r) > 10 and r[0].isdigit()]
    print(f, len(rows))
    for r in rows[-45:]:
        print(r[0], r[4][:40], r[-3], r[-2], r[-1])
PY
