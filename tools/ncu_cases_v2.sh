#!/bin/bash
# ncu --set full captures, second set (halo-mode conv on / off, BatchNorm backward apply, RoIAlign backward, sampler)
# usage (on the GPU box): bash tools/ncu_cases_v2.sh
mkdir -p gpurun_out
N="ncu --set full --clock-control none --import-source on"
env HD_HALO=1 $N -k regex:conv_gemm_kernel -s 3 -c 1 -f -o gpurun_out/r2_ncu_fwd3x3_64_160_halo python tools/conv_microbench.py fwd3x3_64_64_160_relu 1 > gpurun_out/n1.log 2>&1
env HD_HALO=0 $N -k regex:conv_gemm_kernel -s 3 -c 1 -f -o gpurun_out/r2_ncu_fwd3x3_64_160_classic python tools/conv_microbench.py fwd3x3_64_64_160_relu 1 > gpurun_out/n2.log 2>&1
env HD_HALO=1 $N -k regex:conv_gemm_kernel -s 3 -c 1 -f -o gpurun_out/r2_ncu_fwd3x3_128_32_256x320_halo python tools/conv_microbench.py fwd3x3_128_32_256x320 1 > gpurun_out/n3.log 2>&1
$N -k regex:bn_bwd_apply_kernel -s 2 -c 1 -f -o gpurun_out/r2_ncu_bn_bwd_apply_128x160x64 python tools/bn_microbench.py > gpurun_out/n4.log 2>&1
$N -k regex:roi_align_bwd_sep_kernel -s 2 -c 1 -f -o gpurun_out/r2_ncu_roi_align_bwd_sep python tools/roi_bwd_bench.py > gpurun_out/n5.log 2>&1
$N -k regex:samp_select_kernel -s 2 -c 1 -f -o gpurun_out/r2_ncu_samp_select python tools/roi_bwd_bench.py > gpurun_out/n6.log 2>&1
ls -la gpurun_out/r2_ncu_*.ncu-rep
