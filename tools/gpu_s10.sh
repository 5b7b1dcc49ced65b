#!/bin/bash
# GPU session 10: team stagger sweep (variants/lib_stag*.so), n=20 G=5
mkdir -p gpurun_out
for ns in 0 1800 2600 5000; do
  echo "== stagger $ns ns"
  DIFFQC_B200_LIB=$PWD/variants/lib_stag$ns.so SKIP_GENERIC=1 ENGINES=1 KG=5 timeout 300 python tools/quick_bench.py 2>&1 | tee -a gpurun_out/s10_stagger.txt
done
echo "== default (3500)"
SKIP_GENERIC=1 ENGINES=1 KG=5 timeout 300 python tools/quick_bench.py 2>&1 | tee -a gpurun_out/s10_stagger.txt
