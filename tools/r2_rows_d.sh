#!/bin/bash
# ncu --set full of the row-lane solve at full population (private factors)
python tools/one_batch.py case2869pegase 4096 1 > gpurun_out/plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"k_solve_rows" -s 40 -c 2 -o gpurun_out/prof_r2_rows -f python tools/one_batch.py case2869pegase 4096 1 > gpurun_out/ncu_r2_rows.log 2>&1
tail -2 gpurun_out/ncu_r2_rows.log
