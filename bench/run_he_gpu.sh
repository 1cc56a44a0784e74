set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_he.py -x -q 2>&1 | tail -15 > gpurun_out/t_he.log; cat gpurun_out/t_he.log
timeout 300 python bench/he_micro.py > gpurun_out/he_micro_v2.log 2>&1; tail -4 gpurun_out/he_micro_v2.log
HEFL_HE_V1=1 timeout 300 python bench/he_micro.py > gpurun_out/he_micro_v1.log 2>&1; tail -4 gpurun_out/he_micro_v1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'encrypt2_kernel|ntt2_kernel|decrypt2_kernel' -s 4 -c 4 -o gpurun_out/prof_he2_n8192 python bench/he_prof.py n8192_l4 600 > gpurun_out/ncu_he2.log 2>&1; tail -3 gpurun_out/ncu_he2.log
