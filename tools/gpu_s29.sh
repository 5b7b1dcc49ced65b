#!/bin/bash
# GPU session 29: tile size of the fused rotation pass (2^12 / 2^11 / 2^10 amplitudes), n = 28 on one GPU
mkdir -p gpurun_out
for tb in 12 11 10; do
  DQ_SLICE_TILE_BITS=$tb N=28 STEPS=2 FUSED=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_slice --csv --log-file gpurun_out/s29_t$tb.csv python tools/dist_state_run.py > gpurun_out/s29_t$tb.log 2>&1
  python - $tb <<'PY'
import csv, sys
tb = sys.argv[1]
rows = [r for r in csv.reader(open("gpurun_out/s29_t%s.csv" % tb)) if len(r) > 10 and r[0].isdigit()]
step = [r for r in rows if "rx_tile" in r[4] or "phase" in r[4]]
n_steps = sum(1 for r in step if "phase" in r[4])
tot = sum(float(r[-1]) for r in step[-(len(step) // max(1, n_steps)):]) / 1e6
print("tile bits", tb, "launches in last step", len(step) // max(1, n_steps), "ms in last step %.3f" % tot, [round(float(r[-1]) / 1e6, 2) for r in step[-(len(step) // max(1, n_steps)):]])
PY
  grep "energy_after" gpurun_out/s29_t$tb.log
done
