#!/bin/bash
# Round 2, GPU call AV: row-aligned (padded) edge-list streams -- 16-byte loads of 8 (column, value) pairs in the row walk.
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_edgelist_gpu.py tests/test_spmm_gpu.py tests/test_optin_kernels_gpu.py -x -q -m gpu -p no:cacheprovider ) > gpurun_out/r2av_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2av_pytest.log
timeout 600 python bench.py --configs pcrystk02_n8,pcrystk02_n16,pcrystk02_n32,pcrystk02_n64 --no-cpu-baseline > gpurun_out/r2av_bench.json 2> gpurun_out/r2av_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2av_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2av_bench.json'))
print('headline us', round(d['ms_per_step']*1e3,3), 'k20', round(d['run']['k_step_graphs']['ms_per_step']*1e3,3), 'batched us', d['batched']['ms_per_spmm']*1e3, d['batched']['bit_exact_every_triple'], 'e2e', round(d['e2e']['ms_per_step']*1e3,1), 'pipelined', round(d['e2e']['pipelined']['ms_per_step']*1e3,1), 'bit', d['parity']['bit_exact_all_ranks'])
for k,x in d['configs'].items(): print('   ',k, x['ms'], x['frac'], x['bit_exact'], x['kernel'][:75])
PY
