#!/bin/bash
# per-step timeline of the two solves (is the side branch the critical path?)
export HELM_B200_LIB=$PWD/helmpy_b200/libhelm_b200_tl.so
for s in "HELM_B200_ROWS_SOLVE=0" "HELM_B200_ROWS_SOLVE=1"; do
  echo "setting: $s"
  env $s python tools/step_times.py case2869pegase 16384 4096 2>&1 | sed -n 1,3p
  env $s python tools/step_times.py case2869pegase 16384 4096 2>&1 | sed -n 60,100p
done
