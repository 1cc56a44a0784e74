#!/bin/bash
# Medical-CNN artefacts after a kernel change (about 4 GPU-minutes):
#   /usr/local/graft/bin/gpurun --timeout 1200 -- 'bash bench/run_medcnn_refresh.sh'
# then here: python bench/make_profiles.py && python bench/roofline.py
set -u
mkdir -p gpurun_out
T="timeout 280"
$T python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_final.log
timeout 240 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke_final.log
$T python bench/nn_micro.py      > gpurun_out/nn_micro.log 2>&1
$T python bench/wgrad0_micro.py  > gpurun_out/wgrad0_micro.log 2>&1
$T python bench/fwd0_micro.py    > gpurun_out/fwd0_micro.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -f"
$T $NCU -c 1 -k regex:pair_fwd_kernel  -o gpurun_out/prof_fwd0_pair   python bench/scratch/one_step.py 3 > /dev/null 2>&1
$T $NCU -c 1 -k regex:wgrad0_mma       -o gpurun_out/prof_wgrad0_mma  python bench/scratch/one_step.py 3 > /dev/null 2>&1
$T ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 75 -c 45 --csv \
   --log-file gpurun_out/launches_v7.csv python bench/scratch/one_step.py 4 > /dev/null 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err
tail -3 gpurun_out/pytest_gpu_final.log; tail -2 gpurun_out/smoke_final.log; tail -1 gpurun_out/wgrad0_micro.log; tail -1 gpurun_out/fwd0_micro.log; cut -c1-260 gpurun_out/bench_1gpu.json
