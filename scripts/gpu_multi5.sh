#!/bin/bash
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -4
show() { python -c "
import json,sys
t=open(sys.argv[1]).read().strip().splitlines()
assert len(t)==1, ('stdout must be ONE line', len(t))
d=json.loads(t[0]);print('  ms',round(d['ms_per_step'],4),'GF',round(d['value'],1),d['config']['launch'][:14],'|',d['config']['partition'][:80])" $1; }
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 200 --warmup 3 $EXTRA > gpurun_out/m5.json 2> gpurun_out/m5.err; echo "nasa x$N rc=$?"; grep -v "^\*\|OMP_NUM\|NCCL version" gpurun_out/m5.err | tail -3 | cut -c1-300; show gpurun_out/m5.json
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 200 --warmup 3 --peer-mode memops > gpurun_out/m5b.json 2> gpurun_out/m5b.err; echo "nasa memops x$N rc=$?"; show gpurun_out/m5b.json
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --workload pcrystk02 --steps 200 --warmup 3 > gpurun_out/m5c.json 2> gpurun_out/m5c.err; echo "pcrystk02 fused x$N rc=$?"; show gpurun_out/m5c.json
