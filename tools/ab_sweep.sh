export HELM_RESIDENT=4096
run() { python tools/sweep.py case2869pegase 16384 1 2>&1 | grep case | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['cases_per_s'], d['ms'], d['steps'])"; }
echo "depth2 split"; run
echo "depth2 fused"; HELM_B200_FUSED_EPS=1 run
echo "depth1 split"; HELM_B200_SIDE_DEPTH=1 run
echo "depth3 split"; HELM_B200_SIDE_DEPTH=3 run
echo "depth2 split noprio"; HELM_B200_SIDE_PRIO=0 run
echo "depth2 split R=4144"; HELM_RESIDENT=4144 run
echo "small"; python tools/sweep.py case2869pegase 1,64,512,4096 1 2>&1 | grep case | cut -c1-140
