#!/bin/bash
# GPU session 28 (round-1 closing run at HEAD): full GPU suite, smoke, bench (both arms), slice launch list at n = 28
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/s28_tests.log 2>&1
tail -5 gpurun_out/s28_tests.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1 | tee gpurun_out/s28_smoke.log
N=28 STEPS=2 FUSED=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_slice --csv --log-file gpurun_out/s28_fused.csv python tools/dist_state_run.py > gpurun_out/s28_fused.log 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/s28_fused.csv")) if len(r) > 10 and r[0].isdigit()]
for r in rows[-10:]:
    print(r[0], r[4][:44], r[-1])
PY
grep "energy_after\|norm2_after" gpurun_out/s28_fused.log
timeout 600 python bench.py > gpurun_out/s28_bench.json 2> gpurun_out/s28_bench.err
cat gpurun_out/s28_bench.json | cut -c1-400
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/s28_bench_ref.json 2> gpurun_out/s28_bench_ref.err
cat gpurun_out/s28_bench_ref.json | cut -c1-200
