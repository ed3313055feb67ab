timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r.csv python bench.py --steps 2 --warmup 1 > gpurun_out/b.log 2>&1
timeout 600 python bench.py > gpurun_out/bench_r.json 2> gpurun_out/bench_r.err; tail -1 gpurun_out/bench_r.json | cut -c1-330
