#!/bin/bash
# Round 2, GPU call BK: 4-lane / 8-lane groups in 128-thread blocks (32 / 16 rows per block) against 256-thread blocks.
mkdir -p gpurun_out
for v in default g4t128 g8t128; do
  lib=""; [ $v != default ] && lib=$PWD/sextans_b200/variants/libsextans_b200_$v.so
  SX_LIBRARY_PATH=$lib timeout 600 python bench.py --configs pcrystk02_n16,pcrystk02_n32 --no-cpu-baseline --no-pipelined-e2e > gpurun_out/r2bk_$v.json 2> gpurun_out/r2bk_$v.err
  python - $v <<'PY'
import json,sys
v=sys.argv[1]
try:
    d=json.load(open(f'gpurun_out/r2bk_{v}.json'))
    print(v, 'headline us', round(d['ms_per_step']*1e3,3), d['roofline']['kernel'][:58], 'batched', d['batched']['ms_per_spmm']*1e3, 'e2e', round(d['e2e']['ms_per_step']*1e3,1), d['parity']['bit_exact_all_ranks'])
    for k,x in d['configs'].items(): print('   ',k, x['ms'], x['bit_exact'], x['kernel'][:70])
except Exception as e: print(v,'failed',e)
PY
done
