#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err; echo "bench rc=$?"
timeout 900 python -m pytest tests/test_gpu_train.py -q --timeout 300 -p no:cacheprovider > gpurun_out/r2w_train_tests.log 2>&1; tail -2 gpurun_out/r2w_train_tests.log
