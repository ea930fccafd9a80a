#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests/test_modules_gpu.py tests/test_unet_layers_gpu.py -m gpu -q -s --tb=short -p no:cacheprovider 2>&1 | grep -v Warning | tail -80 | tee gpurun_out/mod_all.log
