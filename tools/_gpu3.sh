set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2c_topo.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_sharded_check.py > gpurun_out/r2c_sharded_check.log 2>&1; echo "sharded_check rc=$?" >> gpurun_out/r2c_status.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r2c_bench_2gpu.json 2> gpurun_out/r2c_bench_2gpu.err; echo "bench2 rc=$?" >> gpurun_out/r2c_status.txt
cat gpurun_out/r2c_status.txt; tail -5 gpurun_out/r2c_sharded_check.log; tail -5 gpurun_out/r2c_bench_2gpu.err
