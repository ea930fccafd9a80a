#!/bin/bash
# Run the GPU kernel parity tests in separate processes (a deadlocked kernel only loses its own group).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
for grp in "test_conv_fwd" "test_conv_dgrad" "test_conv_wgrad" "test_pack or test_stem" "test_batchnorm or test_bn" "test_maxpool or test_upsample or test_layout or test_resize or test_regulariser"; do
  name=$(echo "$grp" | tr ' ' '_')
  echo "=== $grp ==="
  timeout -s KILL 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "$grp" --tb=short -p no:cacheprovider 2>&1 | tail -40 | tee gpurun_out/kern_$name.log
done
