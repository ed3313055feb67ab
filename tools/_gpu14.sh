mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2n_pytest_all.log 2>&1; echo "pytest_all rc=$?"
timeout 600 python tools/microbench.py --only fusedce,catalog --iters 10 > gpurun_out/r2n_microbench.jsonl 2> gpurun_out/r2n_microbench.err; echo "mb rc=$?"
tail -25 gpurun_out/r2n_pytest_all.log
