#!/bin/bash
# Round 2, GPU call B: the edge-list kernel (variant 5) -- parity, then A/B against variant 3 on the SuiteSparse configs.
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_edgelist_gpu.py tests/test_spmm_gpu.py -x -q -p no:cacheprovider ) > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2b_pytest.log
show() { python -c "import json,sys; d=json.load(open(sys.argv[1])); print('  %-46s kernel us %7.2f  frac %.3f  warm us %6.2f  iso us %6.1f  e2e us %6.1f  %s' % (sys.argv[2], d['ms_per_step']*1e3, d['roofline']['frac'], d['single_copy_back_to_back']['ms_per_step']*1e3, d['isolated_cold_launch']['ms']*1e3, d['e2e']['ms_per_step']*1e3, d['roofline']['kernel'][:40]))" "$1" "$2" 2>/dev/null || echo "  $2: FAILED"; }
run() { tag=$1; shift; python bench.py --no-cpu-baseline --steps 100 "$@" > gpurun_out/r2b_$tag.json 2> gpurun_out/r2b_$tag.err; show gpurun_out/r2b_$tag.json "$tag"; }
run nasa_v5
run nasa_v5_nopdl --pdl 0
run nasa_v5_nopf --prefetch 0
run nasa_v5_nopdl_nopf --pdl 0 --prefetch 0
run nasa_v3 --kernel 3
run nasa_v3_pdl --kernel 3 --pdl 1
run nasa_v1 --kernel 1
for n in 8 16 32 64; do
  run pcr_n${n}_v5 --workload pcrystk02 --ncols $n
  run pcr_n${n}_v5_nopdl --workload pcrystk02 --ncols $n --pdl 0
  run pcr_n${n}_v3 --workload pcrystk02 --ncols $n --kernel 3
done
run pcr_n16_v3_wr64 --workload pcrystk02 --ncols 16 --kernel 3 --window-rows 64
run nasa_f32_v5 --dtype f32
run fem100_f64_v5 --workload fem --band 100 --steps 20
run fem100_f32_v5 --workload fem --band 100 --dtype f32 --steps 20
run fem2000_f64_v5 --workload fem --steps 20
run fem2000_f64_v2 --workload fem --steps 20 --kernel 2
for tool in racecheck synccheck; do
  timeout 400 compute-sanitizer --tool $tool --num-cuda-barriers 65536 --error-exitcode 9 python -m pytest tests/test_edgelist_gpu.py tests/test_spmm_gpu.py -q -p no:cacheprovider -k "small_golden or every_kernel_variant or 1000-1000-8 or 70-64-4" > gpurun_out/r2b_sanitizer_$tool.log 2>&1
  echo "compute-sanitizer $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/r2b_sanitizer_$tool.log | tail -3
done
