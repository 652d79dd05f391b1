#!/bin/bash
# Round 2, GPU call BH: edge-list blocks cut by nonzeros with at most 1.25 / 1.5 sweeps per block (SX_EDGE_BALANCE=4 / 5) against blocks of ROWS rows.
mkdir -p gpurun_out
for b in 0 4 5; do
  SX_EDGE_BALANCE=$b timeout 600 python bench.py --configs pcrystk02_n8,pcrystk02_n16,pcrystk02_n32,pcrystk02_n64 --no-cpu-baseline --no-pipelined-e2e --batch 0 > gpurun_out/r2bh_b$b.json 2> gpurun_out/r2bh_b$b.err
  python - $b <<'PY'
import json,sys
b=sys.argv[1]
try:
    d=json.load(open(f'gpurun_out/r2bh_b{b}.json'))
    print('balance',b,'headline us', round(d['ms_per_step']*1e3,3), d['roofline']['kernel'][:64], d['parity']['bit_exact_all_ranks'])
    for k,x in d['configs'].items(): print('   ',k, x['ms'], x['bit_exact'], x['kernel'][:75])
except Exception as e: print(b,'failed',e)
PY
done
