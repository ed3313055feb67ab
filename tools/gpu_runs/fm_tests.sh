#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fm.py -q --timeout 300 -p no:cacheprovider > gpurun_out/fm_tests.log 2>&1
grep -E "Error|^FAILED|passed|failed|^E  " gpurun_out/fm_tests.log | head -40
