#!/bin/bash
# Round 2, GPU call AP: fatter edge-list blocks (two sweeps of the lane groups): SX_EDGE_BALANCE=2 (one wave where two sweeps
# make it one) and 3 (always 2 x ROWS rows) against the default.
mkdir -p gpurun_out
for b in 0 2 3; do
  SX_EDGE_BALANCE=$b timeout 600 python bench.py --configs pcrystk02_n8,pcrystk02_n16,pcrystk02_n32,pcrystk02_n64 --no-cpu-baseline --batch 0 > gpurun_out/r2ap_b$b.json 2> gpurun_out/r2ap_b$b.err
  python - $b <<'PY'
import json,sys
b=sys.argv[1]
try:
    d=json.load(open(f'gpurun_out/r2ap_b{b}.json'))
    print('balance',b,'headline us', round(d['ms_per_step']*1e3,3), d['roofline']['kernel'][:60], 'e2e', round(d['e2e']['ms_per_step']*1e3,1), d['parity']['bit_exact_all_ranks'])
    for k,x in d['configs'].items(): print('   ',k, x['ms'], x['frac'], x['bit_exact'], x['kernel'][:75])
except Exception as e: print(b,'failed',e)
PY
done
