#!/bin/bash
# Round 2, GPU call H: nnz-balanced edge-list blocks with row sweeps -- parity, timeline, bench with all configs.
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_edgelist_gpu.py tests/test_spmm_gpu.py -x -q -p no:cacheprovider ) > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2h_pytest.log
SX_LIBRARY_PATH=$PWD/sextans_b200/variants/libsextans_b200_trace.so python scripts/edge_trace.py > gpurun_out/r2h_trace_nasa.txt 2>&1; tail -12 gpurun_out/r2h_trace_nasa.txt
SX_LIBRARY_PATH=$PWD/sextans_b200/variants/libsextans_b200_trace.so python scripts/edge_trace.py --workload pcrystk02 --copies 30 > gpurun_out/r2h_trace_pcr.txt 2>&1; tail -6 gpurun_out/r2h_trace_pcr.txt
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2h_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2h_bench.json'))
print('headline us', d['ms_per_step']*1e3, 'frac', d['roofline']['frac'], 'e2e us', d['e2e']['ms_per_step']*1e3, d['run']['timed'], d['roofline']['kernel'])
print('parity', d['parity'])
for k,v in d['configs'].items(): print(k, v)
PY
python bench.py --gpus 1 --steps 20 --warmup 5 --configs none --pdl 0 --no-cpu-baseline > gpurun_out/r2h_bench_nopdl.json 2>/dev/null; python -c "import json; d=json.load(open('gpurun_out/r2h_bench_nopdl.json')); print('nopdl headline us', d['ms_per_step']*1e3)"
