#!/bin/bash
# GPU session 18 (round-1 closing evidence, HEAD build with evict_last bulk stores): full GPU suite, smoke, bench line,
# ncu launch list of the bench command, one full ncu capture of the chained pass kernel.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/s18_tests.log 2>&1
tail -5 gpurun_out/s18_tests.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/s18_smoke.log
timeout 600 python bench.py > gpurun_out/s18_bench.json 2> gpurun_out/s18_bench.err
cat gpurun_out/s18_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/s18_bench_ref.json 2> gpurun_out/s18_bench_ref.err
cat gpurun_out/s18_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s18_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/s18_bench_under_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/s18_launches.csv > gpurun_out/s18_launch_summary.txt 2>&1; tail -12 gpurun_out/s18_launch_summary.txt
G=5 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fused_passes -s 1 -c 1 -f -o gpurun_out/s18_prof python tools/profile_case.py > gpurun_out/s18_ncu.log 2>&1
tail -2 gpurun_out/s18_ncu.log
