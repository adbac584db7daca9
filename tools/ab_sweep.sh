export HELM_RESIDENT=4096
run() { python tools/sweep.py case2869pegase 16384 1 2>&1 | grep case | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['cases_per_s'], d['ms'], d['steps'])"; }
echo "two chains, solve prio"; run
echo "two chains, no prio"; HELM_B200_NO_SOLVE_PRIO=1 run
echo "one chain"; HELM_B200_NO_CHAINS=1 run
echo "two chains prio R=8192"; HELM_RESIDENT=8192 run
echo "two chains prio R=5600"; HELM_RESIDENT=5600 run
