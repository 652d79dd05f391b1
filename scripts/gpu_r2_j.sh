#!/bin/bash
# Round 2, GPU call J: full GPU suite; staged kernel with / without the software-pipelined batch loop on C5 and C4.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -x -q -m gpu -p no:cacheprovider ) > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2j_pytest.log
one() { tag=$1; shift; "$@" > gpurun_out/r2j_$tag.json 2> gpurun_out/r2j_$tag.err; python -c "import json; d=json.load(open('gpurun_out/r2j_$tag.json')); print('%-28s kernel ms %.4f frac %.3f  %s' % ('$tag', d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel'][:80]))" || tail -3 gpurun_out/r2j_$tag.err; }
one c5_pipe python bench.py --workload powerlaw --configs none --no-cpu-baseline --steps 10
SX_LIBRARY_PATH=$PWD/sextans_b200/variants/libsextans_b200_nopipe.so one c5_nopipe python bench.py --workload powerlaw --configs none --no-cpu-baseline --steps 10
one c5_pipe_nopf python bench.py --workload powerlaw --configs none --no-cpu-baseline --steps 10 --prefetch 0
one c5_pipe_item256 python bench.py --workload powerlaw --configs none --no-cpu-baseline --steps 10 --item-nnz 256
one c5_pipe_item1024 python bench.py --workload powerlaw --configs none --no-cpu-baseline --steps 10 --item-nnz 1024
one c4_pipe python bench.py --workload uniform --configs none --no-cpu-baseline --steps 10
SX_LIBRARY_PATH=$PWD/sextans_b200/variants/libsextans_b200_nopipe.so one c4_nopipe python bench.py --workload uniform --configs none --no-cpu-baseline --steps 10
one c4_pipe_pf python bench.py --workload uniform --configs none --no-cpu-baseline --steps 10 --prefetch 1
one pcr64_v2_pipe python bench.py --workload pcrystk02 --ncols 64 --kernel 2 --configs none --no-cpu-baseline
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2j_bench.json'))
print('headline us', d['ms_per_step']*1e3, 'frac', d['roofline']['frac'], 'e2e us', d['e2e']['ms_per_step']*1e3, d['roofline']['kernel'])
for k,v in d['configs'].items(): print(k, v['ms'], v['frac'], v['parity'], v['kernel'][:70])
PY
