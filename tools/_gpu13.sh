mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_catalog.py tests/test_gpu_topk.py tests/test_gpu_options.py tests/test_gpu_golden.py -x -q > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?"
timeout 600 python tools/microbench.py --only fusedce,catalog,scores --iters 10 > gpurun_out/r2m_microbench.jsonl 2> gpurun_out/r2m_microbench.err; echo "mb rc=$?"
tail -15 gpurun_out/r2m_pytest.log
