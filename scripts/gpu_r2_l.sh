#!/bin/bash
# Round 2, GPU call L (2 GPUs): N>1 parity tests, bench at N=2 with the fused push and the strong-scaling configs.
mkdir -p gpurun_out
( timeout 500 python -m pytest tests/test_multi_gpu.py -x -q -p no:cacheprovider ) > gpurun_out/r2l_pytest_multi.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2l_pytest_multi.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 --configs none ) > gpurun_out/r2l_bench2.json 2> gpurun_out/r2l_bench2.err; echo "bench2 rc=$?"; tail -5 gpurun_out/r2l_bench2.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2l_bench2.json'))
    print('N=2 headline us', d['ms_per_step']*1e3, 'value', d['value'], 'e2e us', d['e2e']['ms_per_step']*1e3, d['e2e']['path'])
    print(d['run']['timed']); print(d['parity'])
    for k,v in d['configs'].items(): print(k, v)
except Exception as ex: print('no line', ex)
PY
