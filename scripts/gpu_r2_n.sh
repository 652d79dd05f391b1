#!/bin/bash
mkdir -p gpurun_out
SX_LIBRARY_PATH=$PWD/sextans_b200/variants/libsextans_b200_trace.so timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 scripts/edge_trace_n2.py > gpurun_out/r2n_trace_n2.txt 2>&1; echo rc=$?; grep -v "^W\|NCCL\|^$" gpurun_out/r2n_trace_n2.txt | tail -50
