#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/r2w_bench_2gpu.json 2> gpurun_out/r2w_bench_2gpu.err; echo "bench2 rc=$?"
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --workload dlrm-train --steps 50 --warmup 5 > gpurun_out/r2w_bench_train.json 2> gpurun_out/r2w_bench_train.err; echo "train rc=$?"
