#!/bin/bash
# Multi-GPU evidence run: bash bench/run_mgpu.sh <N>  (under gpurun --gpus N)
N=${1:-2}
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_multiproc.py -x -q -m "gpu and multigpu" 2>&1 | tail -8 > gpurun_out/t_mgpu_$N.log; cat gpurun_out/t_mgpu_$N.log
timeout 600 $TR --master-port 29701 bench/allreduce_sweep.py --out gpurun_out/allreduce_sweep_${N}gpu.json > gpurun_out/sweep_$N.log 2>&1; tail -3 gpurun_out/sweep_$N.log
timeout 600 $TR --master-port 29702 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; cat gpurun_out/bench_${N}gpu.json | cut -c1-900
timeout 600 $TR --master-port 29703 bench.py --gpus $N --steps 3 --warmup 3 --model resnet18 --he-preset n8192_l4 --local-epochs 1 --steps-per-epoch 5 --val-steps 1 --no-own-baseline > gpurun_out/bench_resnet18_${N}gpu.json 2> gpurun_out/bench_resnet18_${N}gpu.err; cat gpurun_out/bench_resnet18_${N}gpu.json | cut -c1-1200
