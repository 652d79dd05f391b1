#!/bin/bash
mkdir -p gpurun_out
show() { python -c "
import json,sys;d=json.load(open(sys.argv[1]));print('  ms',round(d['ms_per_step'],4),'1copy',round(d['single_copy_back_to_back']['ms_per_step'],4),'iso',round(d['isolated_cold_launch']['ms'],4),'GF',round(d['value'],1),'frac',round(d['roofline']['frac'],4),'e2e_ms',round(d['e2e']['ms_per_step'],4),d['e2e']['path'][:10],d['roofline']['kernel'])" $1; }
for wlk in nasa4704 pcrystk02; do
 for k in 1 2 3; do
  timeout 300 python bench.py --workload $wlk --steps 200 --kernel $k --no-cpu-baseline > gpurun_out/p7_${wlk}_k$k.json 2> gpurun_out/p7_${wlk}_k$k.err; echo "$wlk k=$k rc=$?"; tail -2 gpurun_out/p7_${wlk}_k$k.err; show gpurun_out/p7_${wlk}_k$k.json
 done
done
for wlk in uniform powerlaw; do
  timeout 300 python bench.py --workload $wlk --steps 20 --no-cpu-baseline > gpurun_out/p7_${wlk}.json 2> gpurun_out/p7_${wlk}.err; echo "$wlk rc=$?"; tail -2 gpurun_out/p7_${wlk}.err; show gpurun_out/p7_${wlk}.json
done
