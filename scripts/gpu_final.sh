#!/bin/bash
# one-GPU validation as the driver runs it, plus the profile captures kept under profiles/
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -x -q -m gpu ) > gpurun_out/fin_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/fin_pytest.log
( time python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/fin_smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/fin_smoke.log
( time python bench.py --impl reference ) > gpurun_out/fin_bench_ref.json 2> gpurun_out/fin_bench_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/fin_bench_ref.json
( time python bench.py ) > gpurun_out/fin_bench.json 2> gpurun_out/fin_bench.err; echo "bench rc=$?"; cat gpurun_out/fin_bench.json | cut -c1-3500; tail -3 gpurun_out/fin_bench.err
show() { python -c "
import json,sys;d=json.load(open(sys.argv[1]));print('  ms',round(d['ms_per_step'],4),'GF',round(d['value'],1),'frac',round(d['roofline']['frac'],4),'e2e',round(d['e2e']['value'],1),'cpu',round(d['cpu_baseline']['value'],2),d['cpu_baseline']['kind'],d['roofline']['kernel'][:40])" $1; }
for n in 8 16 32 64; do python bench.py --workload pcrystk02 --ncols $n --steps 200 > gpurun_out/fin_bench_pcrystk02_n$n.json 2> /dev/null; echo "bench pcrystk02 N=$n rc=$?"; show gpurun_out/fin_bench_pcrystk02_n$n.json; done
for wlk in uniform powerlaw; do
  python bench.py --workload $wlk --steps 20 > gpurun_out/fin_bench_$wlk.json 2> gpurun_out/fin_bench_$wlk.err; echo "bench $wlk rc=$?"; show gpurun_out/fin_bench_$wlk.json
done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'spmm_|major|pull|flag' -s 190 -c 80 --csv --log-file gpurun_out/fin_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/fin_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spmm_ -s 100 -c 2 -o gpurun_out/fin_prof_nasa python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/fin_ncu_full.log 2>&1
