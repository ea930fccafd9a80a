#!/bin/bash
# ncu --set full captures of representative launches (one kernel each) -> gpurun_out/r2_ncu_*.ncu-rep
# usage (on the GPU box): bash tools/ncu_cases.sh
mkdir -p gpurun_out
N="ncu --set full --clock-control none --import-source on"
run() { # name, env, kernel regex, microbench filter
  env $2 $N -k regex:$3 -s 3 -c 1 -f -o gpurun_out/r2_ncu_$1 python tools/conv_microbench.py $4 1 > gpurun_out/r2_ncu_$1.log 2>&1
}
run fwd3x3_256_32x40_streamk HD_STREAMK=1 conv_gemm_kernel fwd3x3_256_256_32x40
run fwd3x3_256_32x40_classic HD_STREAMK=0 conv_gemm_kernel fwd3x3_256_256_32x40
run fwd3x3_128_64x80_classic HD_STREAMK=0 conv_gemm_kernel fwd3x3_128_128_64x80
run fwd1x1_64_256_160 HD_STREAMK=1 conv_gemm_kernel fwd1x1_64_256_160
run fwd3x3_256_256_160 HD_STREAMK=1 conv_gemm_kernel fwd3x3_256_256_160
run wgrad3x3_256_32x40 HD_STREAMK=1 wgrad_gemm_kernel wgrad3x3_256_256_32x40
run wgrad3x3_64_128x160 HD_STREAMK=1 wgrad_gemm_kernel wgrad3x3_64_64_128x160
ls -la gpurun_out/r2_ncu_*.ncu-rep
