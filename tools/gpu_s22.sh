#!/bin/bash
# GPU session 22: dense tests with the noisy-estimator fixture, other configs, ncu capture of the resident dense kernel
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_dense_gpu.py tests/test_ising_gpu.py -m gpu -x -q ) > gpurun_out/s22_tests.log 2>&1
tail -5 gpurun_out/s22_tests.log
K3=4 timeout 600 python tools/bench_configs.py > gpurun_out/s22_configs.log 2>&1
cp gpurun_out/configs.json gpurun_out/s22_configs.json 2>/dev/null
python - <<'PY'
import json
d = json.load(open("gpurun_out/s22_configs.json"))
print(json.dumps(d["config0_demo_maxcut"]))
print(json.dumps(d["config1_h2_vqe"]))
PY
timeout 300 python tools/dense_small_case.py 2>&1 | tee gpurun_out/s22_dense_small.txt
K=16384 timeout 300 python tools/dense_small_case.py 2>&1 | tee -a gpurun_out/s22_dense_small.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_small -s 2 -c 2 -f -o gpurun_out/s22_small_prof python tools/dense_small_case.py > gpurun_out/s22_ncu.log 2>&1
tail -2 gpurun_out/s22_ncu.log
