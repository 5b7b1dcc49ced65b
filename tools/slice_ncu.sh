#!/bin/bash
# ncu captures of the TMA tile pass kernel inside bench.py --config 4 (one GPU, n=29): launch list of the run, a full capture of a
# plain rotation pass and of a boundary pass ([owed rotations][phase][rotations]).  Usage: gpurun -- 'bash tools/slice_ncu.sh TAG'
TAG=${1:-slice}
O=gpurun_out
CMD="python bench.py --config 4 --steps 4 --warmup 2"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/${TAG}_launches.csv $CMD > $O/${TAG}_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_slice_rx_tma --launch-skip 12 --launch-count 1 -o $O/${TAG}_plain -f $CMD > $O/${TAG}_plain.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_slice_rx_tma --launch-skip 13 --launch-count 1 -o $O/${TAG}_boundary -f $CMD > $O/${TAG}_boundary.log 2>&1
for k in plain boundary; do python tools/ncu_summary.py $O/${TAG}_$k.ncu-rep > $O/${TAG}_${k}_summary.txt 2>&1; head -30 $O/${TAG}_${k}_summary.txt; done
grep -c k_slice $O/${TAG}_launches.csv
