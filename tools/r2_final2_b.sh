#!/bin/bash
# round-2 final measurements (second session), part b (ncu): launch list of the bench command, full captures of the main kernels
CMD="python bench.py --steps 1 --warmup 1 --batch 16384 --no-cpu-baseline --parity 8"
$CMD > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 3000 --csv --log-file gpurun_out/launches_r2_f.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
python tools/one_batch.py case2869pegase 4096 1 > gpurun_out/plain_one.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_solve_rows|k_eps|k_rhs_conv" -s 150 -c 6 -o gpurun_out/prof_r2_f_main -f python tools/one_batch.py case2869pegase 4096 1 > gpurun_out/ncu_main.log 2>&1
echo "main rc=$?"
python tools/one_batch.py case2869pegase 1 1 > gpurun_out/plain_single.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_solve_rows_deep|k_factor_coop" -s 10 -c 3 -o gpurun_out/prof_r2_f_single -f python tools/one_batch.py case2869pegase 1 1 > gpurun_out/ncu_single.log 2>&1
echo "single rc=$?"
