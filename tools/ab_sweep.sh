export HELM_RESIDENT=4096
python tools/sweep.py case2869pegase 16384 1 2>&1 | grep case | cut -c1-160
python tools/sweep.py case2869pegase 16384 1 profile 2>&1 | grep case | cut -c100-500
