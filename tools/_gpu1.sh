set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_lookup_v2.py tests/test_gpu_ops.py -x -q > gpurun_out/r2a_pytest_v2.log 2>&1; echo "pytest_v2 rc=$?" >> gpurun_out/r2a_status.txt
for cfg in "MM_IMMA_V1=1" "MM_IMMA_WARPS=12" "MM_IMMA_WARPS=14" "MM_IMMA_WARPS=16"; do
  echo "== $cfg" >> gpurun_out/r2a_microbench.jsonl
  env $cfg timeout 300 python tools/microbench.py --only fused,fused2,interact --iters 30 >> gpurun_out/r2a_microbench.jsonl 2>> gpurun_out/r2a_microbench.err
done
timeout 300 python bench.py --steps 30 --no-cpu-baseline > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?" >> gpurun_out/r2a_status.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:interact_v2 -s 3 -c 1 -o gpurun_out/r2a_fused_v2 -f python tools/run_kernel.py fused > gpurun_out/r2a_ncu.log 2>&1; echo "ncu rc=$?" >> gpurun_out/r2a_status.txt
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2a_pytest_all.log 2>&1; echo "pytest_all rc=$?" >> gpurun_out/r2a_status.txt
tail -3 gpurun_out/r2a_pytest_v2.log; cat gpurun_out/r2a_status.txt; tail -3 gpurun_out/r2a_pytest_all.log
