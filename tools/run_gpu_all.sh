#!/bin/bash
# kernels + modules + short bench, each in its own process with a hard timeout
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -25 | tee gpurun_out/kern_all.log
timeout -s KILL 900 python -m pytest tests/test_modules_gpu.py tests/test_unet_layers_gpu.py tests/test_tail_batched.py -m gpu -q --tb=short -p no:cacheprovider 2>&1 | grep -v Warning | tail -25 | tee gpurun_out/mod_all.log
HD_PROFILE_DUMP=gpurun_out/launch_profile_events.json timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_short.json 2>gpurun_out/bench_short.err; tail -c 1500 gpurun_out/bench_short.json; tail -5 gpurun_out/bench_short.err
