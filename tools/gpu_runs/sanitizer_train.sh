#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 86 --print-limit 20 python -m pytest tests/test_gpu_train.py -q -x -p no:cacheprovider -k "test_interact_backward or test_sparse_rows_apply or test_bce_head or test_train_step_gradients or (test_dense_wgrad_and_dgrad and (37 or 1000)) or test_dense_apply or test_out_of_range" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/sanitizer_memcheck.log | head -10
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 87 --print-limit 20 python -m pytest tests/test_gpu_train.py -q -x -p no:cacheprovider -k "(test_interact_backward and 64-5-37) or (test_sparse_rows_apply and True-16-sgd) or (test_dense_wgrad_and_dgrad and 128-64-37) or test_train_step_gradients" > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"
grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/sanitizer_racecheck.log | head -10
