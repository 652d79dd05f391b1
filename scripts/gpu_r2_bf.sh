#!/bin/bash
# Round 2, GPU call BF: the one-kernel host call's grid-wide wait polled without __nanosleep (A/B), 1/2/4 column groups.
mkdir -p gpurun_out
MTX=$(python -c "from sextans_b200 import workloads as w; print(w.suitesparse_path('nasa4704'))")
for v in default nosleep; do
  for g in 1 2 4; do
    echo -n "$v: "; if [ $v = default ]; then scripts/micro/e2e_c $MTX 16 $g 2; else LD_PRELOAD=$PWD/sextans_b200/variants/libsextans_b200_nosleep.so scripts/micro/e2e_c $MTX 16 $g 2; fi
  done
done 2>&1 | tee gpurun_out/r2bf_e2e_c.txt
