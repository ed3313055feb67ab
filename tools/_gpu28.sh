#!/bin/bash
# first GPU run of the training-step kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_train.py -q --timeout 300 -p no:cacheprovider > gpurun_out/train_tests.log 2>&1
echo "exit $?" >> gpurun_out/train_tests.log
tail -60 gpurun_out/train_tests.log
