#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
show() { python -c "
import json,sys;d=json.load(open(sys.argv[1]));print('  ms',round(d['ms_per_step'],4),'1copy',round(d['single_copy_back_to_back']['ms_per_step'],4),'iso',round(d['isolated_cold_launch']['ms'],4),'GF',round(d['value'],1),'frac',round(d['roofline']['frac'],4),'e2e_ms',round(d['e2e']['ms_per_step'],4),d['e2e']['path'][:10],d['roofline']['kernel'])" $1; }
timeout 300 python bench.py --steps 200 --no-cpu-baseline > gpurun_out/p11_nasa.json 2> gpurun_out/p11_nasa.err; echo "nasa rc=$?"; tail -2 gpurun_out/p11_nasa.err; show gpurun_out/p11_nasa.json
for n in 8 16 32 64; do for k in 1 2; do timeout 300 python bench.py --workload pcrystk02 --ncols $n --kernel $k --steps 200 --no-cpu-baseline > gpurun_out/p11_pcrystk02_n${n}_k$k.json 2>/dev/null; echo "pcrystk02 N=$n k=$k"; show gpurun_out/p11_pcrystk02_n${n}_k$k.json; done; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:sx:: -s 250 -c 120 --csv --log-file gpurun_out/launches11_nasa.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu11.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches11_nasa.csv')) if len(r)>10 and r[0].isdigit()]
print("launch sequence tail:")
for r in rows[-16:]: print("  ", r[4][:70], r[-1], "ns")
PY
