#!/bin/bash
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
show() { python -c "
import json,sys
t=open(sys.argv[1]).read().strip().splitlines()
assert len(t)==1, ('stdout must be ONE line', len(t))
d=json.loads(t[0]);print('  ms',round(d['ms_per_step'],4),'GF',round(d['value'],1),'e2e_ms',round(d['e2e']['ms_per_step'],4),d['config']['launch'][:14],'|',d['config']['partition'][:110])" $1; }
for wlk in nasa4704 pcrystk02; do
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --workload $wlk --steps 200 --warmup 3 > gpurun_out/m4_$wlk.json 2> gpurun_out/m4_$wlk.err; echo "$wlk x$N rc=$?"; grep -v "^\*\|OMP_NUM" gpurun_out/m4_$wlk.err | tail -4 | cut -c1-300; show gpurun_out/m4_$wlk.json
done
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 200 --warmup 3 --no-graph > gpurun_out/m4_eager.json 2> gpurun_out/m4_eager.err; echo "nasa eager x$N rc=$?"; show gpurun_out/m4_eager.json
