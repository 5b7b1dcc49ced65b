#!/bin/bash
# GPU session 21: full GPU suite with the resident dense engine and vectorised host pulse tables; other configs again
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/s21_tests.log 2>&1
tail -6 gpurun_out/s21_tests.log
K3=4 timeout 600 python tools/bench_configs.py > gpurun_out/s21_configs.log 2>&1
tail -3 gpurun_out/s21_configs.log
cp gpurun_out/configs.json gpurun_out/s21_configs.json 2>/dev/null
python - <<'PY'
import json
d = json.load(open("gpurun_out/s21_configs.json"))
print(json.dumps(d["config0_demo_maxcut"]))
print(json.dumps(d["config1_h2_vqe"]))
PY
