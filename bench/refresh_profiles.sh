#!/bin/bash
# One GPU call that regenerates every single-GPU artefact under gpurun_out/ (≈3 GPU-minutes on a B200):
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash bench/refresh_profiles.sh'
# then, back on the CPU box:
#   python bench/make_profiles.py && python bench/roofline.py && python bench/make_sass.py
# (multi-GPU artefacts: torchrun bench/allreduce_sweep.py and bench.py --gpus N, see README).
set -u
mkdir -p gpurun_out
T="timeout 250"
$T python bench/nn_micro.py      > gpurun_out/nn_micro.log 2>&1
$T python bench/he_micro.py      > gpurun_out/he_micro.log 2>&1
$T python bench/resnet_micro.py  > gpurun_out/resnet_micro.log 2>&1
$T python bench/graph_step.py    > gpurun_out/graph_step.log 2>&1
$T python bench/timeline.py      > gpurun_out/timeline.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -c 1 -f"
$T $NCU -k regex:tap_gemm_kernel -o gpurun_out/prof_fwd0_v4 python bench/one_step.py 3 > /dev/null 2>&1
$T $NCU -k regex:wgrad0_gather   -o gpurun_out/prof_gather python bench/one_step.py 3 > /dev/null 2>&1
$T $NCU -k regex:head_cluster    -o gpurun_out/prof_head_cluster python bench/one_step.py 3 > /dev/null 2>&1
$T $NCU -k regex:wgrad_kernel --launch-skip 4 -o gpurun_out/prof_wgrad1 python bench/one_step.py 3 > /dev/null 2>&1
$T ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 75 -c 45 --csv \
   --log-file gpurun_out/launches_v5.csv python bench/one_step.py 4 > /dev/null 2>&1
$T python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err
tail -2 gpurun_out/graph_step.log; cut -c1-200 gpurun_out/bench_1gpu.json
