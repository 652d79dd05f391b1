#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_edgelist_gpu.py tests/test_spmm_gpu.py -x -q -p no:cacheprovider ) > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2e_pytest.log
show() { python -c "import json,sys; d=json.load(open(sys.argv[1])); print('  %-46s kernel us %7.2f  frac %.3f  warm us %6.2f  iso us %6.1f  e2e us %6.1f  %s' % (sys.argv[2], d['ms_per_step']*1e3, d['roofline']['frac'], d['single_copy_back_to_back']['ms_per_step']*1e3, d['isolated_cold_launch']['ms']*1e3, d['e2e']['ms_per_step']*1e3, d['roofline']['kernel'][:40]))" "$1" "$2" 2>/dev/null || { echo "  $2: FAILED"; tail -3 gpurun_out/r2e_$2.err; }; }
run() { tag=$1; shift; python bench.py --no-cpu-baseline --steps 100 "$@" > gpurun_out/r2e_$tag.json 2> gpurun_out/r2e_$tag.err; show gpurun_out/r2e_$tag.json "$tag"; }
run nasa_v5
run nasa_v5_nopdl --pdl 0
run nasa_v5_nopf --prefetch 0
run nasa_v5_nopdl_nopf --pdl 0 --prefetch 0
run nasa_v5_warm --no-flush
run nasa_v3_pdl_warm --no-flush --kernel 3 --pdl 1
for n in 8 16 32 64; do
  run pcr_n${n}_v5 --workload pcrystk02 --ncols $n
  run pcr_n${n}_v5_nopdl --workload pcrystk02 --ncols $n --pdl 0
done
