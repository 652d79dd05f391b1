#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
show() { python -c "
import json,sys;d=json.load(open(sys.argv[1]));print('  ms',round(d['ms_per_step'],4),'1copy',round(d['single_copy_back_to_back']['ms_per_step'],4),'GF',round(d['value'],1),'frac',round(d['roofline']['frac'],4),'e2e_ms',round(d['e2e']['ms_per_step'],4),d['roofline']['kernel'][:60])" $1; }
for k in 3 1; do
  timeout 300 python bench.py --steps 200 --kernel $k --no-cpu-baseline > gpurun_out/w_nasa_k$k.json 2> gpurun_out/w.err; echo "nasa k=$k rc=$?"; tail -2 gpurun_out/w.err; show gpurun_out/w_nasa_k$k.json
  for n in 8 16 32; do timeout 300 python bench.py --workload pcrystk02 --ncols $n --kernel $k --steps 200 --no-cpu-baseline > gpurun_out/w_pc_n${n}_k$k.json 2> gpurun_out/w.err; echo "pcrystk02 N=$n k=$k"; tail -2 gpurun_out/w.err; show gpurun_out/w_pc_n${n}_k$k.json; done
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:spmm_window -s 100 -c 1 -o gpurun_out/prof_window_nasa python bench.py --kernel 3 --steps 5 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_w.log 2>&1
