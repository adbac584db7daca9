run() { python tools/sweep.py case2869pegase $1 1 2>&1 | grep case | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['cases_per_s'], d['ms'], d['steps'])"; }
for R in 4096 8192 8800 9100 12288; do echo "R=$R B=32768"; HELM_RESIDENT=$R run 32768; done
echo "R=4096 B=16384"; HELM_RESIDENT=4096 run 16384
echo "R=8800 B=16384"; HELM_RESIDENT=8800 run 16384
