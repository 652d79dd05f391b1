#!/bin/bash
# Round 2, GPU call C: launch-floor microbenchmark, the cases call B could not run, ncu launch lists.
mkdir -p gpurun_out
scripts/micro/launch_floor > gpurun_out/r2c_launch_floor.txt 2>&1; cat gpurun_out/r2c_launch_floor.txt
( timeout 600 python -m pytest tests/test_edgelist_gpu.py tests/test_spmm_gpu.py -x -q -p no:cacheprovider ) > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2c_pytest.log
show() { python -c "import json,sys; d=json.load(open(sys.argv[1])); print('  %-46s kernel us %7.2f  frac %.3f  warm us %6.2f  iso us %6.1f  e2e us %6.1f  %s' % (sys.argv[2], d['ms_per_step']*1e3, d['roofline']['frac'], d['single_copy_back_to_back']['ms_per_step']*1e3, d['isolated_cold_launch']['ms']*1e3, d['e2e']['ms_per_step']*1e3, d['roofline']['kernel'][:40]))" "$1" "$2" 2>/dev/null || { echo "  $2: FAILED"; tail -3 gpurun_out/r2c_$2.err; }; }
run() { tag=$1; shift; python bench.py --no-cpu-baseline --steps 100 "$@" > gpurun_out/r2c_$tag.json 2> gpurun_out/r2c_$tag.err; show gpurun_out/r2c_$tag.json "$tag"; }
for n in 32 64; do
  run pcr_n${n}_v5 --workload pcrystk02 --ncols $n
  run pcr_n${n}_v5_nopdl --workload pcrystk02 --ncols $n --pdl 0
done
run pcr_n16_v5_pdl_nopf --workload pcrystk02 --ncols 16 --prefetch 0
run pcr_n8_v5_pdl_nopf --workload pcrystk02 --ncols 8 --prefetch 0
run fem100_f64_v5 --workload fem --band 100 --steps 20
run fem100_f64_v5_nopdl --workload fem --band 100 --steps 20 --pdl 0
run fem100_f32_v5 --workload fem --band 100 --dtype f32 --steps 20
run fem2000_f64_v5 --workload fem --steps 20
for k in 5 3; do
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:spmm -c 40 --csv --log-file gpurun_out/r2c_ncu_nasa_k$k.csv python bench.py --no-cpu-baseline --steps 4 --warmup 3 --kernel $k --no-graph > /dev/null 2> gpurun_out/r2c_ncu_nasa_k$k.err
  echo "ncu k=$k rc=$?"; tail -4 gpurun_out/r2c_ncu_nasa_k$k.csv | cut -c1-300
done
