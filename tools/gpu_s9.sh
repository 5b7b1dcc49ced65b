#!/bin/bash
# GPU session 9: chained groups (one launch per sample) -- parity, quick timing, bench
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_ising_gpu.py -q -x ) > gpurun_out/s9_tests.log 2>&1
tail -5 gpurun_out/s9_tests.log
SKIP_GENERIC=1 ENGINES=1 KG=4,5,6,8 timeout 300 python tools/quick_bench.py 2>&1 | tee gpurun_out/s9_quick.txt
N=16 B=2 SKIP_GENERIC=1 ENGINES=1 KG=40,80 timeout 300 python tools/quick_bench.py 2>&1 | tee gpurun_out/s9_quick16.txt
timeout 600 python bench.py > gpurun_out/s9_bench.json 2> gpurun_out/s9_bench.err
python -c "
import json; d=json.load(open('gpurun_out/s9_bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['roofline']['launches_timed'], d['linear_estimator']['value'])"
