#!/bin/bash
# First GPU call of the next round: what could not be run at the end of round 1.
#   1. the driver's own GPU test command
#   2. the experimental paths' parity tests (SX_OPT_HOST_FUSED)
#   3. e2e of the default bench with and without the fused host path
#   4. staged-kernel tile size (SX_STAGE_KB) on the C5-like probe, with the L2 prefetch
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests -x -q -m gpu -p no:cacheprovider ) > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2a_pytest.log
SX_TEST_EXPERIMENTAL=1 timeout 200 python -m pytest tests/test_experimental_gpu.py -q -p no:cacheprovider > gpurun_out/r2a_experimental.log 2>&1; echo "experimental rc=$?"; tail -15 gpurun_out/r2a_experimental.log
for flag in "" "--host-fused" "--pdl"; do
  python bench.py --no-cpu-baseline $flag > gpurun_out/r2a_bench_nasa$flag.json 2> gpurun_out/r2a_bench_nasa$flag.err; echo "bench nasa4704 $flag rc=$?"
  python -c "import json,sys; d=json.load(open(sys.argv[1])); print('  kernel us', round(d['ms_per_step']*1e3,2), 'e2e us', round(d['e2e']['ms_per_step']*1e3,1), d['e2e']['path'])" gpurun_out/r2a_bench_nasa$flag.json
  python bench.py --workload pcrystk02 --steps 200 --no-cpu-baseline $flag > gpurun_out/r2a_bench_pcrystk02$flag.json 2> /dev/null
  python -c "import json,sys; d=json.load(open(sys.argv[1])); print('  kernel us', round(d['ms_per_step']*1e3,2), 'e2e us', round(d['e2e']['ms_per_step']*1e3,1), d['e2e']['path'])" gpurun_out/r2a_bench_pcrystk02$flag.json
done
for n in 8 16 32; do for wr in 0 64 128; do
  python bench.py --workload pcrystk02 --ncols $n --steps 200 --no-cpu-baseline --window-rows $wr > gpurun_out/r2a_pcrystk02_n${n}_wr$wr.json 2>/dev/null
  python -c "import json,sys; d=json.load(open(sys.argv[1])); print('pcrystk02 N=$n window-rows=$wr: kernel us', round(d['ms_per_step']*1e3,2), d['roofline']['kernel'][:50])" gpurun_out/r2a_pcrystk02_n${n}_wr$wr.json
done; done
for dt in f32 f64; do for wr in 0 64 128; do
  python bench.py --workload fem --band 100 --dtype $dt --kernel 3 --steps 20 --no-cpu-baseline --window-rows $wr > gpurun_out/r2a_fem_band100_${dt}_wr$wr.json 2>/dev/null
  python -c "import json,sys; d=json.load(open(sys.argv[1])); print('fem band=100 $dt window-rows=$wr: ms', round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],3), d['roofline']['kernel'][:50])" gpurun_out/r2a_fem_band100_${dt}_wr$wr.json
done; done
for dt in f32 f64; do for sl in 1 2; do
  python bench.py --workload fem --band 100 --dtype $dt --kernel 4 --slide $sl --steps 20 --no-cpu-baseline > gpurun_out/r2a_fem_band100_${dt}_slide$sl.json 2>/dev/null
  python -c "import json,sys; d=json.load(open(sys.argv[1])); print('fem band=100 $dt slide=$sl: ms', round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],3), d['roofline']['kernel'][:50])" gpurun_out/r2a_fem_band100_${dt}_slide$sl.json
done; done
for wlk in nasa4704 pcrystk02 uniform powerlaw; do
  python bench.py --workload $wlk --steps 20 --no-cpu-baseline --autotune --slide 1 > gpurun_out/r2a_autotune_$wlk.json 2>/dev/null
  python -c "import json,sys; d=json.load(open(sys.argv[1])); print('autotune $wlk: ms', round(d['ms_per_step'],5), d['roofline']['kernel'][:60])" gpurun_out/r2a_autotune_$wlk.json
done
for kb in 28 56 84; do
  SX_STAGE_KB=$kb PROBE_SET=0:-1 timeout 100 python scripts/probe_windows.py 2>&1 | tail -1 | sed "s/^/stage_kb=$kb /"
done | tee gpurun_out/r2a_stage_kb.log
# alternative build of the staged kernel: 16 gathers in flight per lane, 2 blocks per SM (fp64).
# Build it BEFORE the gpurun call, here on the CPU:
#   scripts/build_variant.sh umax16 "-DSX_STAGED_UMAX=16 -DSX_STAGED_MINBLOCKS_F64=2"
if [ -f sextans_b200/variants/libsextans_b200_umax16.so ]; then
  SX_LIBRARY_PATH=$PWD/sextans_b200/variants/libsextans_b200_umax16.so PROBE_SET=0:-1 timeout 100 python scripts/probe_windows.py 2>&1 | tail -1 | sed "s/^/umax16 /" | tee gpurun_out/r2a_umax16.log
fi
# compute-sanitizer on the small configs: memcheck and racecheck of the
# golden / canned-run tests, default paths and (gated) experimental ones
for tool in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_spmm_gpu.py -q -p no:cacheprovider -k "small_golden or config2 or every_kernel_variant" > gpurun_out/r2a_sanitizer_$tool.log 2>&1
  echo "compute-sanitizer $tool rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2a_sanitizer_$tool.log | tail -3
done
SX_TEST_EXPERIMENTAL=1 timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_experimental_gpu.py -q -p no:cacheprovider -k "64-64-4 or 70-64-4 or 1000" > gpurun_out/r2a_sanitizer_experimental.log 2>&1
echo "compute-sanitizer experimental rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2a_sanitizer_experimental.log | tail -3

