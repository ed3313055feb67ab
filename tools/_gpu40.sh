#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -q --timeout 300 -p no:cacheprovider > gpurun_out/train_tests.log 2>&1
grep -E "AssertionError|^FAILED|passed|failed|Error" gpurun_out/train_tests.log | head -20
