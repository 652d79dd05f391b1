#!/bin/bash
# ncu --set full of the staged kernel as it now runs on the C5-like matrix (trimmed loop + L2 prefetch)
mkdir -p gpurun_out
PROBE_SET=0:-1 PROBE_REPS=1 timeout 60 ncu --set full --clock-control none --import-source on \
  -k regex:spmm_staged -s 2 -c 1 -o gpurun_out/r1_c5_final_prof python scripts/probe_windows.py > gpurun_out/r1_c5_final_prof.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/r1_c5_final_prof.log
timeout 40 python bench.py --workload pcrystk02 --ncols 64 --steps 200 --no-cpu-baseline > gpurun_out/r1_final2_bench_pcrystk02_n64.json 2>/dev/null; echo "bench exit $?"; cut -c1-200 gpurun_out/r1_final2_bench_pcrystk02_n64.json
