#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
show() { python -c "
import json,sys;d=json.load(open(sys.argv[1]));print('  ms',round(d['ms_per_step'],4),'GF',round(d['value'],1),'frac',round(d['roofline']['frac'],4),'e2e_ms',round(d['e2e']['ms_per_step'],3),d['roofline']['kernel'])" $1; }
for t in 0 4; do
  timeout 600 python bench.py --workload fem --tiles $t --steps 20 --no-cpu-baseline > gpurun_out/p16_fem_t$t.json 2> gpurun_out/p16_fem_t$t.err; echo "fem tiles=$t rc=$?"; tail -2 gpurun_out/p16_fem_t$t.err; show gpurun_out/p16_fem_t$t.json
done
for n in 8 32 64; do
  timeout 600 python bench.py --workload fem --tiles 4 --ncols $n --steps 20 --no-cpu-baseline > gpurun_out/p16_fem_t4_n$n.json 2> /dev/null; echo "fem tiles=4 N=$n"; show gpurun_out/p16_fem_t4_n$n.json
  timeout 600 python bench.py --workload fem --tiles 0 --ncols $n --steps 20 --no-cpu-baseline > gpurun_out/p16_fem_t0_n$n.json 2> /dev/null; echo "fem tiles=0 N=$n"; show gpurun_out/p16_fem_t0_n$n.json
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'spmm_' -s 4 -c 1 -o gpurun_out/prof16_fem_tiles python bench.py --workload fem --tiles 4 --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu16a.log 2>&1
