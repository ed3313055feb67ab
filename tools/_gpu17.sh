mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lookup_v2.py tests/test_gpu_models.py tests/test_gpu_golden.py tests/test_gpu_options.py tests/test_gpu_mlp_tc.py -x -q > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?"
timeout 300 python tools/microbench.py --only fused2 --iters 30 > gpurun_out/r2q_microbench.jsonl 2> gpurun_out/r2q_microbench.err; echo "mb rc=$?"
timeout 600 python bench.py --steps 30 --workload dlrm > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err; echo "bench rc=$?"
MM_TABLE_MIRROR=0 timeout 600 python bench.py --steps 30 --workload dlrm --no-cpu-baseline > gpurun_out/r2q_bench_nomirror.json 2> gpurun_out/r2q_bench_nomirror.err; echo "bench2 rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:interact_v2 -s 3 -c 1 -o gpurun_out/r2q_fused_operand -f python tools/run_kernel.py fused_operand > gpurun_out/r2q_ncu.log 2>&1; echo "ncu rc=$?"
tail -8 gpurun_out/r2q_pytest.log
