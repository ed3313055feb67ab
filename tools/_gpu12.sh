mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2l_pytest_all.log 2>&1; echo "pytest_all rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:interact_v2 -s 3 -c 1 -o gpurun_out/r2l_fused_v2 -f python tools/run_kernel.py fused > gpurun_out/r2l_ncu.log 2>&1; echo "ncu rc=$?"
timeout 600 python bench.py --steps 30 > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err; echo "bench rc=$?"
tail -5 gpurun_out/r2l_pytest_all.log
