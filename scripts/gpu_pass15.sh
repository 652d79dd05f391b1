#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu.log
show() { python -c "
import json,sys;d=json.load(open(sys.argv[1]));print('  ms',round(d['ms_per_step'],4),'GF',round(d['value'],1),'frac',round(d['roofline']['frac'],4),'e2e_ms',round(d['e2e']['ms_per_step'],3),d['roofline']['kernel'])" $1; }
for t in 0 4 2 6; do
  timeout 600 python bench.py --workload fem --tiles $t --steps 20 --no-cpu-baseline > gpurun_out/p15_fem_t$t.json 2> gpurun_out/p15_fem_t$t.err; echo "fem tiles=$t rc=$?"; tail -2 gpurun_out/p15_fem_t$t.err; show gpurun_out/p15_fem_t$t.json
done
for t in 0 4; do
  timeout 300 python bench.py --workload pcrystk02 --dtype f64 --tiles $t --steps 200 --no-cpu-baseline > gpurun_out/p15_pc_t$t.json 2> gpurun_out/p15_pc_t$t.err; echo "pcrystk02 f64 tiles=$t rc=$?"; tail -2 gpurun_out/p15_pc_t$t.err; show gpurun_out/p15_pc_t$t.json
  timeout 300 python bench.py --tiles $t --steps 200 --no-cpu-baseline > gpurun_out/p15_nasa_t$t.json 2> gpurun_out/p15_nasa_t$t.err; echo "nasa tiles=$t rc=$?"; tail -2 gpurun_out/p15_nasa_t$t.err; show gpurun_out/p15_nasa_t$t.json
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'spmm_' -s 4 -c 2 -o gpurun_out/prof15_fem_tiles python bench.py --workload fem --tiles 4 --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu15a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'spmm_' -s 3 -c 1 -o gpurun_out/prof15_fem_csr python bench.py --workload fem --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu15b.log 2>&1
