#!/bin/bash
# multi-GPU pass on however many GPUs the box has: tests, then the bench per workload
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
[ -n "$SKIP_TESTS" ] || timeout 400 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -8
show() { python -c "
import json,sys
t=open(sys.argv[1]).read().strip().splitlines()
assert len(t)==1, ('stdout must be ONE line', len(t))
d=json.loads(t[0]);print('  ms',round(d['ms_per_step'],4),'GF',round(d['value'],1),'e2e_ms',round(d['e2e']['ms_per_step'],4),d['config']['launch'][:14],'|',d['config']['partition'][:110])" $1; }
for wlk in ${WORKLOADS:-nasa4704 pcrystk02 uniform powerlaw}; do
  steps=200; case $wlk in uniform|powerlaw) steps=20;; esac
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --workload $wlk --steps $steps --warmup 3 > gpurun_out/multi${N}_$wlk.json 2> gpurun_out/multi${N}_$wlk.err; echo "$wlk x$N rc=$?"; grep -v "^\*\|OMP_NUM" gpurun_out/multi${N}_$wlk.err | tail -3 | cut -c1-300; show gpurun_out/multi${N}_$wlk.json
done
timeout 200 ./sextans_b200/sextans /tmp/sextans_b200_fixtures/pcrystk02.mtx 64 20 --gpus $N --json 2>&1 | tail -7
