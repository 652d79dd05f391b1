#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
show() { python -c "
import json,sys;d=json.load(open(sys.argv[1]));print('  ms',round(d['ms_per_step'],4),'1copy',round(d['single_copy_back_to_back']['ms_per_step'],4),'iso',round(d['isolated_cold_launch']['ms'],4),'GF',round(d['value'],1),'frac',round(d['roofline']['frac'],4),'e2e_ms',round(d['e2e']['ms_per_step'],4),d['e2e']['path'][:10],d['roofline']['kernel'])" $1; }
timeout 300 python bench.py --steps 200 --no-cpu-baseline > gpurun_out/p10_nasa.json 2> gpurun_out/p10_nasa.err; echo "nasa rc=$?"; tail -2 gpurun_out/p10_nasa.err; show gpurun_out/p10_nasa.json
for n in 8 16 32 64; do timeout 300 python bench.py --workload pcrystk02 --ncols $n --steps 200 --no-cpu-baseline > gpurun_out/p10_pcrystk02_n$n.json 2>/dev/null; echo "pcrystk02 N=$n"; show gpurun_out/p10_pcrystk02_n$n.json; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches10_nasa.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu10.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches10_nasa.csv')) if len(r)>10 and r[0].isdigit()]
from collections import defaultdict
d=defaultdict(list)
for r in rows: d[r[4][:60]].append(float(r[-1]))
for k,v in d.items(): print(f"{len(v):4d} x {sum(v)/len(v)/1000:8.2f} us  {k}")
PY
