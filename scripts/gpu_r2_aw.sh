#!/bin/bash
# Round 2, GPU call AW: the row-aligned walk -- register / in-flight variants (default: 8 B pieces in flight, look-ahead of the
# next chunk's columns; al3: the same held to 80 registers; halves: B pieces four at a time; halves3).
mkdir -p gpurun_out
for v in default al3 halves halves3; do
  lib=""; [ $v != default ] && lib=$PWD/sextans_b200/variants/libsextans_b200_$v.so
  SX_LIBRARY_PATH=$lib timeout 600 python bench.py --configs pcrystk02_n8,pcrystk02_n16,pcrystk02_n32,pcrystk02_n64 --no-cpu-baseline --no-pipelined-e2e > gpurun_out/r2aw_$v.json 2> gpurun_out/r2aw_$v.err
  python - $v <<'PY'
import json,sys
v=sys.argv[1]
try:
    d=json.load(open(f'gpurun_out/r2aw_{v}.json'))
    print(v, 'headline us', round(d['ms_per_step']*1e3,3), 'k20', round(d['run']['k_step_graphs']['ms_per_step']*1e3,3), 'batched us', d['batched']['ms_per_spmm']*1e3, 'e2e', round(d['e2e']['ms_per_step']*1e3,1), 'bit', d['parity']['bit_exact_all_ranks'], d['batched']['bit_exact_every_triple'])
    print('   ', {k: x['ms'] for k,x in d['configs'].items()}, all(x['bit_exact'] for x in d['configs'].values()))
except Exception as e: print(v,'failed',e)
PY
done
