mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2v_pytest_all.log 2>&1; echo "pytest_all rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2v_smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err; echo "bench rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 40 --csv --log-file gpurun_out/r2v_launches.csv python bench.py --steps 4 --warmup 3 --workload dlrm --no-cpu-baseline > gpurun_out/r2v_under_ncu.log 2>&1; echo "launches rc=$?"
tail -4 gpurun_out/r2v_pytest_all.log; tail -2 gpurun_out/r2v_smoke.log
