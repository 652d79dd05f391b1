#!/bin/bash
mkdir -p gpurun_out
timeout 600 ./scripts/micro/gather_bench > gpurun_out/gather_bench.txt 2>&1; echo "micro rc=$?"; cat gpurun_out/gather_bench.txt
python scripts/exp_panels.py 1 2>&1 | tail -6
