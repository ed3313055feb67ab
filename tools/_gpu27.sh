mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 160 -c 36 --csv --log-file gpurun_out/r2v_launches.csv python bench.py --steps 6 --warmup 3 --workload dlrm --no-cpu-baseline > gpurun_out/r2v_under_ncu.log 2>&1; echo "launches rc=$?"
