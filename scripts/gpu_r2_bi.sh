#!/bin/bash
# Round 2, GPU call BI: 2-lane groups (32-byte dense rows) in 128-thread blocks (64 rows per block) against 256-thread blocks (128 rows).
mkdir -p gpurun_out
for v in default g2t128; do
  lib=""; [ $v != default ] && lib=$PWD/sextans_b200/variants/libsextans_b200_$v.so
  for cfg in "pcrystk02 8 f32" "pcrystk02 4 f64" "nasa4704 8 f32" "nasa4704 4 f64"; do set -- $cfg
    SX_LIBRARY_PATH=$lib timeout 300 python bench.py --workload $1 --ncols $2 --dtype $3 --configs none --no-cpu-baseline --no-pipelined-e2e --batch 0 --min-region-ms 20 > gpurun_out/r2bi_tmp.json 2>/dev/null
    python -c "import json; d=json.load(open('gpurun_out/r2bi_tmp.json')); print('$v $1 N=$2 $3:', round(d['ms_per_step']*1e3,3), 'us', d['parity']['bit_exact_all_ranks'], d['roofline']['kernel'][:60])"
  done
done
