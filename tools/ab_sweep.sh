export HELM_RESIDENT=4096
run() { python tools/sweep.py case2869pegase 16384 1 2>&1 | grep case | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['cases_per_s'], d['ms'], d['steps'])"; }
echo "shared lu"; run
echo "no shared lu"; HELM_B200_NO_SHARED_LU=1 run
echo "shared R=4480"; HELM_RESIDENT=4480 run
echo "small"; python tools/sweep.py case2869pegase 1,64,512,4096 1 2>&1 | grep case | cut -c1-140
