#!/bin/bash
# Round 2, GPU call AF (2 GPUs): the host-facing call on every rank of the row-block path (sx_spmm_staged_B_*), cooperative
# launch of the one-kernel call on/off, N=2 bench line.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_edgelist_gpu.py -x -q -m gpu -p no:cacheprovider ) > gpurun_out/r2af_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2af_pytest.log
MTX=$(python -c "from sextans_b200 import workloads as w; print(w.suitesparse_path('nasa4704'))")
( for i in 1 2; do SX_HOST_COOP=1 scripts/micro/e2e_c $MTX 16 0 2; SX_HOST_COOP=0 scripts/micro/e2e_c $MTX 16 0 2; done ) 2>&1 | tee gpurun_out/r2af_e2e_c.txt
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2af_bench2.json 2> gpurun_out/r2af_bench2.err; echo "bench2 rc=$?"; tail -3 gpurun_out/r2af_bench2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2af_bench2.json').read().strip().splitlines()[-1])
print('N=2 headline us', d['ms_per_step']*1e3, 'value', d['value'], 'e2e us', d['e2e']['ms_per_step']*1e3, d['e2e']['path'], 'parity', d['parity']['bit_exact_all_ranks'])
print(d['run']['timed']); print(d['run'].get('k_step_graphs'))
for k,v in d['configs'].items(): print(k, {a:b for a,b in v.items() if a in ('ms_kernel','ms_step','parity_all_ranks','bit_exact_all_ranks','nnz_imbalance')}, (v.get('pipelined') or {}).get('ms_step'))
PY
