#!/bin/bash
# multi-GPU: weak-scaling bench line + strong scaling of config 3 on N GPUs (second session of round 2)
N=${1:-2}
P=$((29500 + N))
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 3 --warmup 2 > gpurun_out/bench_r2_f_${N}gpu.json 2> gpurun_out/bench_r2_f_${N}gpu.err
echo "weak rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P+20)) bench.py --strong --gpus $N --steps 5 --warmup 2 > gpurun_out/strong_r2_f_${N}gpu.json 2> gpurun_out/strong_r2_f_${N}gpu.err
echo "strong rc=$?"
tail -c 400 gpurun_out/strong_r2_f_${N}gpu.json
