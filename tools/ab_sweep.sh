export HELM_RESIDENT=4096
run() { python tools/sweep.py case2869pegase 16384 1 2>&1 | grep case | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['cases_per_s'], d['ms'], d['steps'])"; }
echo "L2 ahead 256"; run
echo "L2 ahead 512"; HELM_B200_LIB=/root/repo/helmpy_b200/libhelm_b200_alt.so run
echo "L2 ahead 128"; HELM_B200_LIB=/root/repo/helmpy_b200/libhelm_b200_alt2.so run
echo "profile 256"; python tools/sweep.py case2869pegase 16384 1 profile 2>&1 | grep case | cut -c100-500
