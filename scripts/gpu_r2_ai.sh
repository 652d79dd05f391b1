#!/bin/bash
# Round 2, GPU call AI (4 GPUs): the multi-GPU worker on 4 ranks (a forwarding rank on the kernels that do not carry the exchange).
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 tests/multi_gpu_worker.py > gpurun_out/r2ai_worker4.log 2>&1; echo "worker4 rc=$?"; grep "^OK" gpurun_out/r2ai_worker4.log; grep -i "assert\|Error" gpurun_out/r2ai_worker4.log | head -5
