mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2p_gpus.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tests/dist_sharded_check.py > gpurun_out/r2p_sharded_check.log 2>&1; echo "sharded_check rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 30 --warmup 5 > gpurun_out/r2p_bench_8gpu.json 2> gpurun_out/r2p_bench_8gpu.err; echo "bench8 rc=$?"
tail -2 gpurun_out/r2p_sharded_check.log; tail -3 gpurun_out/r2p_bench_8gpu.err
