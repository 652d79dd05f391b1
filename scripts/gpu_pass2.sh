#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for wlk in nasa4704 pcrystk02 uniform powerlaw; do
 for k in 0 1; do
  python bench.py --workload $wlk --steps 20 --kernel $k --no-cpu-baseline > gpurun_out/p2_${wlk}_k$k.json 2> gpurun_out/p2_${wlk}_k$k.err; echo "$wlk k=$k rc=$?"
  python -c "
import json;d=json.load(open('gpurun_out/p2_${wlk}_k$k.json'));print('  ms',round(d['ms_per_step'],4),'warm',round(d['steady_state_l2_warm']['ms_per_step'],4),'GF',round(d['value'],1),'frac',round(d['roofline']['frac'],4),'e2e_ms',round(d['e2e']['ms_per_step'],3),d['roofline']['kernel'])"
 done
done
for s in 64 128 512; do
  python bench.py --workload powerlaw --steps 10 --item-nnz $s --no-cpu-baseline > gpurun_out/p2_powerlaw_s$s.json 2>/dev/null
  python -c "
import json;d=json.load(open('gpurun_out/p2_powerlaw_s$s.json'));print('powerlaw item $s ms',round(d['ms_per_step'],4),d['roofline']['kernel'])"
done
python bench.py --workload powerlaw --steps 10 --split 2048 --no-cpu-baseline | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('powerlaw split 2048 ms',round(d['ms_per_step'],4),d['roofline']['kernel'])"
python scripts/exp_panels.py 2>&1 | tail -8
ncu --set full --clock-control none --import-source on -k regex:spmm_ -s 3 -c 3 -o gpurun_out/prof2_powerlaw python bench.py --workload powerlaw --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu2_full2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spmm_items -s 3 -c 1 -o gpurun_out/prof2_uniform python bench.py --workload uniform --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu2_full3.log 2>&1
