#!/bin/bash
# GPU session 8: full GPU test suite and the bench line for the two-team / three-buffer TMA engine
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/s8_tests.log 2>&1
tail -6 gpurun_out/s8_tests.log
timeout 600 python bench.py > gpurun_out/s8_bench.json 2> gpurun_out/s8_bench.err
cat gpurun_out/s8_bench.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
