#!/bin/bash
# Round 2, GPU call M: compute-sanitizer on hardware, ncu DRAM traffic per BASELINE config, ncu --set full of the two
# large-matrix kernels, ncu launch list of the default bench command.
mkdir -p gpurun_out
SEL='small_golden or every_kernel_variant or 1000-1000-8 or 70-64-4 or 333-777-1 or canned or dependent_chain or 999-1200-32'
for tool in memcheck racecheck synccheck; do
  extra=""; [ $tool != memcheck ] && extra="--num-cuda-barriers 65536"
  timeout 900 compute-sanitizer --tool $tool $extra --error-exitcode 9 python -m pytest tests/test_spmm_gpu.py tests/test_edgelist_gpu.py tests/test_optin_kernels_gpu.py -q -p no:cacheprovider -k "$SEL" > gpurun_out/r2m_sanitizer_$tool.log 2>&1
  echo "compute-sanitizer $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/r2m_sanitizer_$tool.log | tail -3
done
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active"
prof() { tag=$1; shift; ncu --metrics $M --clock-control none -k regex:"spmm_(edgelist|staged|window|rows)" -s 12 -c 4 --csv --log-file gpurun_out/r2m_ncu_$tag.csv python bench.py --configs none --no-cpu-baseline --no-graph --steps 3 --warmup 3 --min-region-ms 0.01 "$@" > /dev/null 2> gpurun_out/r2m_ncu_$tag.err; echo "ncu $tag rc=$?"; }
prof nasa4704_n16_f64
for n in 8 16 32 64; do prof pcrystk02_n${n}_f32 --workload pcrystk02 --ncols $n; done
prof uniform_n128_f32 --workload uniform
prof powerlaw_n16_f64 --workload powerlaw
for wl in powerlaw uniform; do
  ncu --set full --import-source on --clock-control none -k regex:spmm_staged -s 6 -c 1 -f -o gpurun_out/r2m_full_$wl python bench.py --workload $wl --configs none --no-cpu-baseline --no-graph --steps 3 --warmup 3 --min-region-ms 0.01 > /dev/null 2> gpurun_out/r2m_full_$wl.err; echo "ncu full $wl rc=$?"
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2m_launches_bench.csv python bench.py --steps 20 --warmup 5 --configs none --no-cpu-baseline --min-region-ms 0.05 > gpurun_out/r2m_launches_bench.json 2> /dev/null; echo "launch list rc=$?"
ls -la gpurun_out/r2m_*
