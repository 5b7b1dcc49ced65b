#!/bin/bash
# Timing experiments on the GPU box: rebuild the fused engine with pieces removed (results are WRONG on
# purpose; only the time matters) to see what each component costs.
cd diffquantum_b200/csrc
for e in 0 1 2 4 3 5 6 7; do
  touch ising_fused.cu
  make EXTRA=-DDQ_EXP=$e > /dev/null 2>&1
  echo "DQ_EXP=$e  (1 no-FP64, 2 no-smem, 4 no-global)"
  (cd ../..; python tools/exp_grid.py 2>&1 | grep '"grid_per_sm": 2, "G": 4\|"grid_per_sm": 1, "G": 4')
done
touch ising_fused.cu; make > /dev/null 2>&1
