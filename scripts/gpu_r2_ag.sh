#!/bin/bash
# Round 2, GPU call AG (2 GPUs): deferred publication of the push (the publish kernel out of the chain of dependent launches).
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_multi_gpu.py -x -q -m gpu -p no:cacheprovider ) > gpurun_out/r2ag_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2ag_pytest.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --configs none > gpurun_out/r2ag_bench2.json 2> gpurun_out/r2ag_bench2.err; echo "bench2 rc=$?"; tail -3 gpurun_out/r2ag_bench2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2ag_bench2.json').read().strip().splitlines()[-1])
print('N=2 headline us', d['ms_per_step']*1e3, 'value', d['value'], 'e2e us', d['e2e']['ms_per_step']*1e3, 'parity', d['parity'])
print(d['run']['timed']); print(d['run'].get('k_step_graphs')); print(d['run']['exchange'][:200])
PY
