#!/bin/bash
# round-2 final measurements, part a (no profiler): tests, bench, reference arm, configs, single case
python -m pytest tests -m gpu -x -q > gpurun_out/r2_final_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2_final_pytest.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2_c_1gpu.json 2> gpurun_out/bench_r2_c_1gpu.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2_c_reference_arm.json 2> gpurun_out/bench_r2_c_reference_arm.err; echo "ref rc=$?"
python tools/single_case.py > gpurun_out/single_case_r2_c.txt 2>&1
python tools/solve_time.py >> gpurun_out/single_case_r2_c.txt 2>&1
rm -f gpurun_out/configs.jsonl
python tests/run_configs.py 1 2 3 4 > gpurun_out/configs_r2_c.log 2>&1; echo "configs rc=$?"
HELM_N1_MAX=4096 python tests/run_configs.py 5 >> gpurun_out/configs_r2_c.log 2>&1; echo "config5 rc=$?"
cp gpurun_out/configs.jsonl gpurun_out/configs_r2_c_1gpu.jsonl
python -c "from helmpy_b200.core import _lib; print(_lib.measure_peaks(), _lib.measure_latencies())" > gpurun_out/peaks_r2_c.txt 2>&1
