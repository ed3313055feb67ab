#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -q --timeout 300 -p no:cacheprovider > gpurun_out/train_tests.log 2>&1
grep -E "AssertionError:|^FAILED|passed|failed|Error" gpurun_out/train_tests.log | head -20
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/train_launches.csv python tools/train_profile.py > gpurun_out/train_profile.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/train_launches.csv')) if len(r)>5 and r[0].isdigit()]
tot=0
for r in rows:
    name=r[4][:60]; v=float(r[-1].replace(',',''));
    tot+=v
    if v>30000: print(f"{v/1000:9.1f} us  {name}")
print("total us", tot/1000, "launches", len(rows))
PY
timeout 600 python bench.py --workload dlrm-train --steps 20 --warmup 3 > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_train.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['eager_phase_ms'])"
