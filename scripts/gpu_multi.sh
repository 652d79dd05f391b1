#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -15
N=$(nvidia-smi -L | wc -l)
for wlk in nasa4704 powerlaw; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --workload $wlk --steps 20 --warmup 3 > gpurun_out/multi_${wlk}_$N.json 2> gpurun_out/multi_${wlk}_$N.err; echo "$wlk x$N rc=$?"; tail -3 gpurun_out/multi_${wlk}_$N.err; cat gpurun_out/multi_${wlk}_$N.json | cut -c1-900
done
./sextans_b200/sextans /tmp/sextans_b200_fixtures/pcrystk02.mtx 64 20 --gpus $N 2>&1 | tail -9
