#!/bin/bash
mkdir -p gpurun_out
SX_LIBRARY_PATH=$PWD/sextans_b200/variants/libsextans_b200_trace.so timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 scripts/edge_trace_n2.py > gpurun_out/r2o_trace_n2.txt 2>&1; echo rc=$?; grep -v "^W\|NCCL\|^$\|\*\*\*\|OMP_NUM" gpurun_out/r2o_trace_n2.txt | awk 'NR<=9 || /rank 1/ || (NR>24 && NR<=32)'
( timeout 500 python -m pytest tests/test_multi_gpu.py -x -q -p no:cacheprovider ) > gpurun_out/r2o_pytest_multi.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2o_pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 --configs none > gpurun_out/r2o_bench2.json 2> gpurun_out/r2o_bench2.err; echo "bench2 rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r2o_bench2.json')); print('N=2 headline us', d['ms_per_step']*1e3, 'value', d['value'], 'e2e us', d['e2e']['ms_per_step']*1e3); print(d['run']['timed']); print(d['parity'])"
