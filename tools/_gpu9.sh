mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_sharded_check.py > gpurun_out/r2i_sharded_check.log 2>&1; echo "sharded_check rc=$?"
for cfg in "MM_IMMA_NBUF=4" "MM_IMMA_NBUF=3" "MM_IMMA_NBUF=4 MM_IMMA_WARPS=6"; do
  echo "== $cfg" >> gpurun_out/r2i_sharded_tune.jsonl
  env $cfg timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 --workload dlrm-sharded >> gpurun_out/r2i_sharded_tune.jsonl 2>> gpurun_out/r2i_sharded_tune.err
done
timeout 300 python tools/probe/fused_peer.py > gpurun_out/r2i_fused_peer.jsonl 2> gpurun_out/r2i_fused_peer.err; echo rc=$?
tail -3 gpurun_out/r2i_sharded_check.log
