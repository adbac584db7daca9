#!/bin/bash
# A/B helper (GPU box): cases/s of one 16384-scenario batch of case2869pegase (4096 resident) for a
# list of environment settings.  Usage: bash tools/ab_sweep.sh "NAME=VALUE ..." "NAME2=VALUE2 ..."
export HELM_RESIDENT=${HELM_RESIDENT:-4096}
run() { python tools/sweep.py case2869pegase ${HELM_AB_BATCH:-16384} 1 2>&1 | grep case | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['cases_per_s'], d['ms'], d['steps'])"; }
echo "default"; run
for setting in "$@"; do echo "$setting"; env $setting bash -c "$(declare -f run); run"; done
