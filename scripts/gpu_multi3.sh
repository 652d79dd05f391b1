#!/bin/bash
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -12
show() { python -c "
import json,sys;d=json.load(open(sys.argv[1]));print('  ms',round(d['ms_per_step'],4),'GF',round(d['value'],1),'e2e_ms',round(d['e2e']['ms_per_step'],4),d['config']['partition'][:150])" $1; }
for wlk in nasa4704 pcrystk02; do
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --workload $wlk --steps 200 --warmup 3 > gpurun_out/m3_$wlk.json 2> gpurun_out/m3_$wlk.err; echo "$wlk x$N rc=$?"; grep -v "^\*\|OMP_NUM" gpurun_out/m3_$wlk.err | tail -4 | cut -c1-300; show gpurun_out/m3_$wlk.json
done
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 200 --warmup 3 --peer-bytes 0 > gpurun_out/m3_nccl.json 2> gpurun_out/m3_nccl.err; echo "nasa nccl x$N rc=$?"; show gpurun_out/m3_nccl.json
