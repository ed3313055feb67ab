set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_lookup_v2.py tests/test_gpu_ops.py tests/test_gpu_models.py -x -q > gpurun_out/r2b_pytest_v2.log 2>&1; echo "pytest_v2 rc=$?" >> gpurun_out/r2b_status.txt
for cfg in "MM_IMMA_K8=0 MM_IMMA_WARPS=16" "MM_IMMA_K8=1 MM_IMMA_WARPS=16" "MM_IMMA_K8=1 MM_IMMA_WARPS=12" "MM_IMMA_K8=0 MM_IMMA_WARPS=12"; do
  echo "== $cfg" >> gpurun_out/r2b_microbench.jsonl
  env $cfg timeout 300 python tools/microbench.py --only fused2,interact --iters 30 >> gpurun_out/r2b_microbench.jsonl 2>> gpurun_out/r2b_microbench.err
done
timeout 600 python bench.py --steps 30 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?" >> gpurun_out/r2b_status.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:interact_v2 -s 3 -c 1 -o gpurun_out/r2b_fused_v2 -f python tools/run_kernel.py fused > gpurun_out/r2b_ncu.log 2>&1; echo "ncu rc=$?" >> gpurun_out/r2b_status.txt
cat gpurun_out/r2b_status.txt; tail -3 gpurun_out/r2b_pytest_v2.log; tail -5 gpurun_out/r2b_bench.err
