#!/bin/bash
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
show() { python -c "
import json,sys
t=open(sys.argv[1]).read().strip().splitlines()
assert len(t)==1, ('stdout must be ONE line', len(t))
d=json.loads(t[0]);print('  ms',round(d['ms_per_step'],4),'GF',round(d['value'],1),d['config']['launch'][:14],'|',d['config']['partition'][:80])" $1; }
for wlk in nasa4704 pcrystk02; do
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --workload $wlk --steps 200 --warmup 3 > gpurun_out/m6_$wlk.json 2> gpurun_out/m6_$wlk.err; echo "$wlk x$N rc=$?"; grep -v "^\*\|OMP_NUM\|NCCL version" gpurun_out/m6_$wlk.err | tail -3 | cut -c1-300; show gpurun_out/m6_$wlk.json
done
