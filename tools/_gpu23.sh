mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mlp_tc.py -x -q > gpurun_out/r2s_pytest.log 2>&1; echo "pytest rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tower_small -s 3 -c 1 -o gpurun_out/r2s_tower_small -f python tools/run_kernel.py bottom > gpurun_out/r2s_ncu.log 2>&1; echo "ncu rc=$?"
timeout 600 python bench.py --steps 30 --workload dlrm --no-cpu-baseline > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err; echo "bench rc=$?"
tail -4 gpurun_out/r2s_pytest.log
