mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mlp_tc.py tests/test_gpu_models.py tests/test_gpu_golden.py tests/test_gpu_lookup_v2.py tests/test_gpu_options.py -x -q > gpurun_out/r2r_pytest.log 2>&1; echo "pytest rc=$?"
timeout 300 python tools/microbench.py --only bottom --iters 30 > gpurun_out/r2r_microbench.jsonl 2>&1
timeout 600 python bench.py --steps 30 --workload dlrm > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err; echo "bench rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 60 --csv --log-file gpurun_out/r2r_launches.csv python bench.py --steps 3 --warmup 3 --workload dlrm --no-cpu-baseline > gpurun_out/r2r_under_ncu.log 2>&1; echo "launches rc=$?"
tail -8 gpurun_out/r2r_pytest.log
