#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
for wlk in nasa4704 pcrystk02 uniform powerlaw; do
 for k in 0 1; do
  timeout 300 python bench.py --workload $wlk --steps 20 --kernel $k --no-cpu-baseline > gpurun_out/p4_${wlk}_k$k.json 2> gpurun_out/p4_${wlk}_k$k.err; echo "$wlk k=$k rc=$?"; tail -2 gpurun_out/p4_${wlk}_k$k.err
  python -c "
import json;d=json.load(open('gpurun_out/p4_${wlk}_k$k.json'));print('  ms',round(d['ms_per_step'],4),'warm',round(d['steady_state_l2_warm']['ms_per_step'],4),'GF',round(d['value'],1),'frac',round(d['roofline']['frac'],4),'e2e_ms',round(d['e2e']['ms_per_step'],3),d['roofline']['kernel'])"
 done
done
for s in 64 128 512; do
  timeout 300 python bench.py --workload powerlaw --steps 10 --item-nnz $s --no-cpu-baseline > gpurun_out/p4_powerlaw_s$s.json 2>/dev/null
  python -c "
import json;d=json.load(open('gpurun_out/p4_powerlaw_s$s.json'));print('powerlaw item $s ms',round(d['ms_per_step'],4),d['roofline']['kernel'])"
done
timeout 300 python scripts/exp_panels.py 0 2>&1 | tail -6
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmm_ -s 3 -c 2 -o gpurun_out/prof4_powerlaw python bench.py --workload powerlaw --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu4_full2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmm_staged -s 3 -c 1 -o gpurun_out/prof4_uniform python bench.py --workload uniform --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu4_full3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmm_staged -s 3 -c 1 -o gpurun_out/prof4_nasa python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu4_full1.log 2>&1
