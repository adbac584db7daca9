#!/bin/bash
export HELM_RESIDENT=4096
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "benchmark_path or large_batch or full_size or continuous or n_minus_1 or islanding or tiled or random_scales" 2>&1 | tail -3
run() {  # label, env
  echo "variant $1"
  env $2 python tools/sweep.py case2869pegase 16384 1 profile 2>&1 | grep '"case"' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms'], d['ms_kernel'], d['conv'])"
  env $2 python tools/sweep.py case2869pegase 16384 1 2>&1 | grep '"case"' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['cases_per_s'], d['ms'], d.get('fc'))"
}
run nocache HELM_B200_NO_FCACHE=1
run cache A=1
