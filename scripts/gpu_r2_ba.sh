#!/bin/bash
# Round 2, GPU call BA: the row-aligned walk held to 64 registers (four resident blocks per SM, 2..8-lane groups) against the default.
mkdir -p gpurun_out
for v in default edge4b; do
  lib=""; [ $v != default ] && lib=$PWD/sextans_b200/variants/libsextans_b200_$v.so
  SX_LIBRARY_PATH=$lib timeout 600 python bench.py --configs pcrystk02_n8,pcrystk02_n16,pcrystk02_n32,pcrystk02_n64 --no-cpu-baseline --no-pipelined-e2e > gpurun_out/r2ba_$v.json 2> gpurun_out/r2ba_$v.err
  python - $v <<'PY'
import json,sys
v=sys.argv[1]
try:
    d=json.load(open(f'gpurun_out/r2ba_{v}.json'))
    print(v, 'headline us', round(d['ms_per_step']*1e3,3), 'k20', round(d['run']['k_step_graphs']['ms_per_step']*1e3,3), 'batched us', d['batched']['ms_per_spmm']*1e3, 'bit', d['parity']['bit_exact_all_ranks'], d['batched']['bit_exact_every_triple'])
    print('   ', {k: x['ms'] for k,x in d['configs'].items()}, all(x['bit_exact'] for x in d['configs'].values()))
except Exception as e: print(v,'failed',e)
PY
done
