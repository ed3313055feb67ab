#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2x_bench_ref.json 2> gpurun_out/r2x_bench_ref.err; echo "ref rc=$?"
