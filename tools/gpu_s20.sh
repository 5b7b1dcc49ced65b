#!/bin/bash
# GPU session 20: resident dense engine (dim <= 16, dense_small.cu): dense parity tests, then the other BASELINE configs
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_dense_gpu.py tests/test_diffqc_shim.py -m gpu -x -q ) > gpurun_out/s20_tests.log 2>&1
tail -15 gpurun_out/s20_tests.log
timeout 600 python tools/bench_configs.py > gpurun_out/s20_configs.log 2>&1
tail -3 gpurun_out/s20_configs.log
cp gpurun_out/configs.json gpurun_out/s20_configs.json 2>/dev/null
python - <<'PY'
import json
d = json.load(open("gpurun_out/s20_configs.json"))
print(json.dumps(d["config0_demo_maxcut"]))
print(json.dumps(d["config1_h2_vqe"]))
print(json.dumps(d["config2_n16"]))
PY
