#!/bin/bash
mkdir -p gpurun_out
for grp in "test_unet_train_forward" "test_unet_eval or test_unet_train_cuda" "test_frozen_backbone" "test_transform" "test_train_step" "test_trainer"; do
  name=$(echo "$grp" | tr ' ' '_')
  echo "=== $grp ==="
  timeout -s KILL 600 python -m pytest tests/test_modules_gpu.py -m gpu -q -s -k "$grp" --tb=short -p no:cacheprovider 2>&1 | tail -60 | tee gpurun_out/mod_$name.log
done
