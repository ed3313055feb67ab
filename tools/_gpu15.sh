mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_sharded_check.py > gpurun_out/r2o_sharded_check.log 2>&1; echo "sharded_check rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r2o_bench_2gpu.json 2> gpurun_out/r2o_bench_2gpu.err; echo "bench2 rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 --impl reference > gpurun_out/r2o_bench_ref.json 2> gpurun_out/r2o_bench_ref.err; echo "ref rc=$?"
tail -2 gpurun_out/r2o_sharded_check.log; tail -3 gpurun_out/r2o_bench_2gpu.err
