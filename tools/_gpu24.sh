mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mlp_tc.py tests/test_gpu_models.py tests/test_gpu_golden.py -x -q > gpurun_out/r2t_pytest.log 2>&1; echo "pytest rc=$?"
timeout 600 python bench.py --steps 30 --workload dlrm --no-cpu-baseline > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err; echo "bench rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:tower_small -s 6 -c 6 --csv --log-file gpurun_out/r2t_launches.csv python bench.py --steps 3 --warmup 3 --workload dlrm --no-cpu-baseline > gpurun_out/r2t_under_ncu.log 2>&1; echo "launches rc=$?"
tail -4 gpurun_out/r2t_pytest.log; grep tower_small gpurun_out/r2t_launches.csv | tail -4 | cut -c1-60,200-
