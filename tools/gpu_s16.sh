#!/bin/bash
# GPU session 16 (round-1 evidence for the two-team TMA engine): full GPU suite, bench line, ncu launch list of the bench
# command, one full ncu capture of the chained pass kernel.
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/s16_tests.log 2>&1
tail -5 gpurun_out/s16_tests.log
timeout 600 python bench.py > gpurun_out/s16_bench.json 2> gpurun_out/s16_bench.err
cat gpurun_out/s16_bench.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s16_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/s16_bench_under_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/s16_launches.csv > gpurun_out/s16_launch_summary.txt 2>&1; tail -12 gpurun_out/s16_launch_summary.txt
G=5 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fused_passes -s 1 -c 1 -f -o gpurun_out/s16_prof python tools/profile_case.py > gpurun_out/s16_ncu.log 2>&1
tail -2 gpurun_out/s16_ncu.log
