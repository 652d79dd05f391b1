#!/bin/bash
# Round 2, GPU call AH (4 GPUs): the push tree with forwarding ranks under deferred publication -- tests and the headline.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_multi_gpu.py -x -q -m gpu -p no:cacheprovider ) > gpurun_out/r2ah_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2ah_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 tests/multi_gpu_worker.py > gpurun_out/r2ah_worker4.log 2>&1; echo "worker4 rc=$?"; grep -c "^OK" gpurun_out/r2ah_worker4.log; tail -3 gpurun_out/r2ah_worker4.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 20 --warmup 5 --configs none > gpurun_out/r2ah_bench4.json 2> gpurun_out/r2ah_bench4.err; echo "bench4 rc=$?"; tail -3 gpurun_out/r2ah_bench4.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2ah_bench4.json').read().strip().splitlines()[-1])
print('N=4 headline us', d['ms_per_step']*1e3, 'value', d['value'], 'e2e us', d['e2e']['ms_per_step']*1e3, 'parity', d['parity'])
print(d['run']['timed']); print(d['run'].get('k_step_graphs'))
PY
