#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py -q --timeout 300 -p no:cacheprovider -k "sparse or training_steps" > gpurun_out/train_tests.log 2>&1
tail -3 gpurun_out/train_tests.log
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/train_launches.csv python tools/train_profile.py > gpurun_out/train_profile.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/train_launches.csv')) if len(r)>5 and r[0].isdigit()]
tot=0
for r in rows:
    name=r[4][:60]; v=float(r[-1].replace(',',''));
    tot+=v
    if v>8000: print(f"{v/1000:9.1f} us  {name}")
print("total us", tot/1000, "launches", len(rows))
PY
timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:"sparse_apply_dense|sparse_scatter_small|interact_bwd|sparse_apply_kernel|dgrad_kernel<8>|wgrad_kernel<4, 2, 4, 4>" -o gpurun_out/train_full python tools/train_profile.py > gpurun_out/train_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
