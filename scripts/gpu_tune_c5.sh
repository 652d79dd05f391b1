#!/bin/bash
mkdir -p gpurun_out
show() { python -c "
import json,sys;d=json.load(open(sys.argv[1]));print('  ms',round(d['ms_per_step'],4),'GF',round(d['value'],1),'frac',round(d['roofline']['frac'],4),d['roofline']['kernel'][40:])" $1; }
for cfg in "--item-nnz 512" "--item-nnz 1024" "--item-nnz 512 --split 1024" "--item-nnz 1024 --split 2048" "--item-nnz 2048 --split 4096"; do
  timeout 300 python bench.py --workload powerlaw --steps 10 --no-cpu-baseline $cfg > gpurun_out/t5.json 2> gpurun_out/t5.err; echo "powerlaw $cfg rc=$?"; tail -1 gpurun_out/t5.err; show gpurun_out/t5.json
done
for cfg in "--item-nnz 512" "--item-nnz 128"; do
  timeout 300 python bench.py --workload uniform --steps 10 --no-cpu-baseline $cfg > gpurun_out/t5.json 2> gpurun_out/t5.err; echo "uniform $cfg rc=$?"; show gpurun_out/t5.json
done
