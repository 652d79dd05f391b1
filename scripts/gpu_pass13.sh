#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python scripts/exp_e2e.py nasa4704 16 f64
python scripts/exp_e2e.py pcrystk02 16 f32
python scripts/exp_e2e.py pcrystk02 64 f32
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'spmm_|major' -s 60 -c 6 --csv --log-file gpurun_out/launches13_e2e.csv python scripts/exp_e2e.py nasa4704 16 f64 > gpurun_out/ncu13.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches13_e2e.csv')) if len(r)>10 and r[0].isdigit()]
for r in rows[:6]: print("  ", r[4][:70], r[-1], "ns")
PY
