#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -q --timeout 300 -p no:cacheprovider > gpurun_out/train_tests.log 2>&1
grep -E "AssertionError|^FAILED|passed|failed|Error" gpurun_out/train_tests.log | head -20
timeout 900 python bench.py --workload dlrm-train --steps 30 --warmup 5 > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_train.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value']); print(d.get('roofline')); print(d.get('cpu_baseline'))"
tail -3 gpurun_out/bench_train.err
