#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
show() { python -c "
import json,sys;d=json.load(open(sys.argv[1]));print('  ms',round(d['ms_per_step'],4),'warm',round(d['single_copy_back_to_back']['ms_per_step'],4),'GF',round(d['value'],1),'frac',round(d['roofline']['frac'],4),'e2e_ms',round(d['e2e']['ms_per_step'],4),d['e2e']['path'][:10],d['roofline']['kernel'])" $1; }
for wlk in nasa4704 pcrystk02; do
 for k in 0 1 2; do
  timeout 300 python bench.py --workload $wlk --steps 30 --kernel $k --no-cpu-baseline > gpurun_out/p6_${wlk}_k$k.json 2> gpurun_out/p6_${wlk}_k$k.err; echo "$wlk k=$k rc=$?"; tail -2 gpurun_out/p6_${wlk}_k$k.err; show gpurun_out/p6_${wlk}_k$k.json
 done
done
for n in 8 32 64; do timeout 300 python bench.py --workload pcrystk02 --ncols $n --steps 30 --no-cpu-baseline > gpurun_out/p6_pcrystk02_n$n.json 2>/dev/null; echo "pcrystk02 N=$n"; show gpurun_out/p6_pcrystk02_n$n.json; done
for wlk in uniform powerlaw; do
  timeout 300 python bench.py --workload $wlk --steps 20 --no-cpu-baseline > gpurun_out/p6_${wlk}.json 2> gpurun_out/p6_${wlk}.err; echo "$wlk rc=$?"; tail -2 gpurun_out/p6_${wlk}.err; show gpurun_out/p6_${wlk}.json
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmm_ -s 3 -c 1 -o gpurun_out/prof6_nasa python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu6_1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmm_ -s 3 -c 1 -o gpurun_out/prof6_pcrystk02 python bench.py --workload pcrystk02 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu6_2.log 2>&1
