#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
show() { python -c "
import json,sys;d=json.load(open(sys.argv[1]));print('  ms',round(d['ms_per_step'],4),'GF',round(d['value'],1),'frac',round(d['roofline']['frac'],4),'e2e_ms',round(d['e2e']['ms_per_step'],4),d['roofline']['kernel'][:70])" $1; }
for cfg in "--kernel 0" "--kernel 2" "--tiles 4"; do
  timeout 400 python bench.py --workload fem --band 100 --steps 20 --no-cpu-baseline $cfg > gpurun_out/w2_fem.json 2> gpurun_out/w2.err; echo "fem band100 $cfg rc=$?"; tail -2 gpurun_out/w2.err; show gpurun_out/w2_fem.json
done
timeout 400 python bench.py --workload fem --band 100 --dtype f32 --steps 20 --no-cpu-baseline > gpurun_out/w2_fem32.json 2> gpurun_out/w2.err; echo "fem band100 f32 auto rc=$?"; tail -2 gpurun_out/w2.err; show gpurun_out/w2_fem32.json
timeout 400 python bench.py --workload fem --band 100 --dtype f32 --kernel 2 --steps 20 --no-cpu-baseline > gpurun_out/w2_fem32k2.json 2> gpurun_out/w2.err; echo "fem band100 f32 k2 rc=$?"; show gpurun_out/w2_fem32k2.json
timeout 400 ncu --set full --clock-control none --import-source on -k regex:spmm_window -s 4 -c 1 -o gpurun_out/prof_window_fem python bench.py --workload fem --band 100 --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_w2.log 2>&1
