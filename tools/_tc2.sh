export MM_INTERACT_TC=1
timeout 300 python -m pytest tests/test_gpu_ops.py tests/test_gpu_golden.py -x -q -k "interact or dlrm or golden" 2>&1 | tail -3
timeout 200 python tools/microbench.py --only interact,fused 2>&1 | tail -2 | cut -c1-200
