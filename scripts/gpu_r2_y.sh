#!/bin/bash
# Round 2, GPU call Y: the one-kernel host call with 1 / 2 / 4 column groups requested ahead (SX_HOST_DEPTH), parity of its tests.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_edgelist_gpu.py -x -q -m gpu -p no:cacheprovider ) > gpurun_out/r2y_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2y_pytest.log
MTX=$(python -c "from sextans_b200 import workloads as w; print(w.suitesparse_path('nasa4704'))")
( for d in 1 2 3; do for g in 2 4 8; do echo -n "depth $d: "; SX_HOST_DEPTH=$d scripts/micro/e2e_c $MTX 16 $g 2; done; done
  scripts/micro/e2e_c $MTX 16 1 2; scripts/micro/e2e_c $MTX 16 1 1 ) 2>&1 | tee gpurun_out/r2y_e2e_c.txt
