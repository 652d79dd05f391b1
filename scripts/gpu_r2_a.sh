#!/bin/bash
# Round 2, GPU call A: validate everything round 1 left unrun (gated tests, PDL, host-fused, tall windows, slide),
# plus compute-sanitizer over the small parity cases.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
( time timeout 300 python -m pytest tests -x -q -m gpu -p no:cacheprovider ) > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2a_pytest.log
SX_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_experimental_gpu.py -q -p no:cacheprovider > gpurun_out/r2a_experimental.log 2>&1; echo "experimental rc=$?"; tail -15 gpurun_out/r2a_experimental.log
for flag in "" "--host-fused" "--pdl"; do
  python bench.py --no-cpu-baseline $flag > gpurun_out/r2a_bench_nasa$flag.json 2> gpurun_out/r2a_bench_nasa$flag.err; echo "bench nasa4704 $flag rc=$?"
  python -c "import json,sys; d=json.load(open(sys.argv[1])); print('  kernel us', round(d['ms_per_step']*1e3,2), 'e2e us', round(d['e2e']['ms_per_step']*1e3,1), d['e2e']['path'])" gpurun_out/r2a_bench_nasa$flag.json
done
for n in 16; do for wr in 0 64 128; do
  python bench.py --workload pcrystk02 --ncols $n --steps 200 --no-cpu-baseline --window-rows $wr > gpurun_out/r2a_pcrystk02_n${n}_wr$wr.json 2>/dev/null
  python -c "import json,sys; d=json.load(open(sys.argv[1])); print('pcrystk02 N=$n window-rows=$wr: kernel us', round(d['ms_per_step']*1e3,2), d['roofline']['kernel'][:50])" gpurun_out/r2a_pcrystk02_n${n}_wr$wr.json
done; done
for dt in f64; do for wr in 0 128; do
  python bench.py --workload fem --band 100 --dtype $dt --kernel 3 --steps 20 --no-cpu-baseline --window-rows $wr > gpurun_out/r2a_fem_band100_${dt}_wr$wr.json 2>/dev/null
  python -c "import json,sys; d=json.load(open(sys.argv[1])); print('fem band=100 $dt window-rows=$wr: ms', round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],3), d['roofline']['kernel'][:50])" gpurun_out/r2a_fem_band100_${dt}_wr$wr.json
done; done
for dt in f32 f64; do for sl in 1; do
  python bench.py --workload fem --band 100 --dtype $dt --kernel 4 --slide $sl --steps 20 --no-cpu-baseline > gpurun_out/r2a_fem_band100_${dt}_slide$sl.json 2>/dev/null
  python -c "import json,sys; d=json.load(open(sys.argv[1])); print('fem band=100 $dt slide=$sl: ms', round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],3), d['roofline']['kernel'][:50])" gpurun_out/r2a_fem_band100_${dt}_slide$sl.json
done; done
for tool in memcheck racecheck synccheck; do
  timeout 500 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_spmm_gpu.py -q -p no:cacheprovider -k "small_golden or config2 or every_kernel_variant" > gpurun_out/r2a_sanitizer_$tool.log 2>&1
  echo "compute-sanitizer $tool rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2a_sanitizer_$tool.log | tail -3
done
