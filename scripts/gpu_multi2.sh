#!/bin/bash
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
show() { python -c "
import json,sys;d=json.load(open(sys.argv[1]));print('  ms',round(d['ms_per_step'],4),'1copy',round(d['single_copy_back_to_back']['ms_per_step'],4),'GF',round(d['value'],1),'frac',round(d['roofline']['frac'],4),'e2e_ms',round(d['e2e']['ms_per_step'],4),d['config']['launch'][:12])" $1; }
for g in "" "--no-graph"; do
timeout 300 python bench.py --steps 200 --no-cpu-baseline $g > gpurun_out/g1.json 2> gpurun_out/g1.err; echo "nasa x1 $g rc=$?"; tail -2 gpurun_out/g1.err; show gpurun_out/g1.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 200 --warmup 3 $g > gpurun_out/g2.json 2> gpurun_out/g2.err; echo "nasa x$N $g rc=$?"; tail -3 gpurun_out/g2.err | cut -c1-300; show gpurun_out/g2.json
done
timeout 300 python bench.py --workload pcrystk02 --steps 200 --no-cpu-baseline > gpurun_out/g1.json 2> gpurun_out/g1.err; echo "pcrystk02 x1 rc=$?"; tail -2 gpurun_out/g1.err; show gpurun_out/g1.json
