#!/bin/bash
export HELM_RESIDENT=4096
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "benchmark_path or large_batch or full_size or continuous or n_minus_1 or tiled" 2>&1 | tail -2
run() {  # label, lib, env
  echo "variant $1"
  env HELM_B200_LIB=$PWD/helmpy_b200/libhelm_b200$2.so $3 python tools/sweep.py case2869pegase 16384 1 profile 2>&1 | grep '"case"' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms'], d['ms_kernel']['solve'], d['ms_kernel']['factor'], d['ms_kernel']['fcache'], d['conv'])"
  env HELM_B200_LIB=$PWD/helmpy_b200/libhelm_b200$2.so $3 python tools/sweep.py case2869pegase 16384 1 2>&1 | grep '"case"' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['cases_per_s'], d['ms'], d.get('fc'))"
}
run cache "" A=1
echo "pool smaller than the number of distinct matrices (recycling): resident 256 of 2048 scenarios"
python - <<'PY'
import sys, os
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import numpy as np
from common import load_case
import helmpy_b200
case = load_case('case1354pegase')
s = np.random.default_rng(3).uniform(0.85, 1.04, 2048)
a = helmpy_b200.helm_batch(case, s, mismatch=1e-8, max_coefficients=40, max_resident=256)
os.environ['HELM_B200_NO_FCACHE'] = '1'
b = helmpy_b200.helm_batch(case, s, mismatch=1e-8, max_coefficients=40, max_resident=256)
print('recycling: stats', a.stats['factor_cache_hits'], a.stats['factorizations'], 'equal ncoef', bool(np.array_equal(a.ncoef, b.ncoef)), 'max dV', float(np.max(np.abs(a.V - b.V))), 'ms', a.stats['ms_total'], b.stats['ms_total'])
PY
