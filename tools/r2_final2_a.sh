#!/bin/bash
# round-2 final measurements (second session), part a (no profiler): smoke, tests, bench, reference arm, single case, configs
python __graft_entry__.py smoke 2>&1 | tail -1
python -m pytest tests -m gpu -x -q > gpurun_out/r2_final3_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2_final2_pytest.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2_f_1gpu.json 2> gpurun_out/bench_r2_f_1gpu.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2_f_reference_arm.json 2> gpurun_out/bench_r2_f_reference_arm.err; echo "ref rc=$?"
python tools/single_case.py > gpurun_out/single_case_r2_f.txt 2>&1
python tools/solve_time.py >> gpurun_out/single_case_r2_f.txt 2>&1
rm -f gpurun_out/configs.jsonl
python tests/run_configs.py 1 2 3 4 > gpurun_out/configs_r2_f.log 2>&1; echo "configs rc=$?"
HELM_N1_MAX=4096 python tests/run_configs.py 5 >> gpurun_out/configs_r2_f.log 2>&1; echo "config5 rc=$?"
cp gpurun_out/configs.jsonl gpurun_out/configs_r2_f_1gpu.jsonl
