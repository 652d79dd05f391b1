#!/bin/bash
mkdir -p gpurun_out
python scripts/exp_e2e.py nasa4704 16 f64
python scripts/exp_e2e.py pcrystk02 16 f32
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'spmm_|major' -s 60 -c 60 --csv --log-file gpurun_out/launches12_e2e.csv python scripts/exp_e2e.py nasa4704 16 f64 > gpurun_out/ncu12.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches12_e2e.csv')) if len(r)>10 and r[0].isdigit()]
for r in rows[:9]: print("  ", r[4][:70], r[-1], "ns")
PY
