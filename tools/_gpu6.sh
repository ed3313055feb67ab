mkdir -p gpurun_out
for cfg in "MM_IMMA_NBUF=2 MM_IMMA_WARPS=16" "MM_IMMA_NBUF=4 MM_IMMA_WARPS=8" "MM_IMMA_NBUF=3 MM_IMMA_WARPS=10" "MM_IMMA_NBUF=4 MM_IMMA_WARPS=6"; do
  echo "== $cfg" >> gpurun_out/r2f_sharded_tune.jsonl
  env $cfg timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 --workload dlrm-sharded >> gpurun_out/r2f_sharded_tune.jsonl 2>> gpurun_out/r2f_sharded_tune.err
done
echo done
