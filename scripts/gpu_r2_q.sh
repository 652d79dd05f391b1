#!/bin/bash
# Round 2, GPU call Q (4 GPUs): the push tree with an inner rank (0 -> {1,2}, 1 -> {3}); bench at N=4 with the strong-scaling configs.
mkdir -p gpurun_out
nvidia-smi -L | head -8
G=${1:-4}
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $G --steps 20 --warmup 5 ) > gpurun_out/r2q_bench$G.json 2> gpurun_out/r2q_bench$G.err; echo "bench$G rc=$?"; tail -5 gpurun_out/r2q_bench$G.err
python - $G <<'PY'
import json, sys
G=sys.argv[1]
try:
    d=json.load(open(f'gpurun_out/r2q_bench{G}.json'))
    print(f'N={G} headline us', d['ms_per_step']*1e3, 'value', d['value'], 'e2e us', d['e2e']['ms_per_step']*1e3, d['e2e']['path'])
    print(d['run']['timed']); print(d['parity'])
    for k,v in d['configs'].items(): print(k, v)
except Exception as ex: print('no line', ex)
PY
