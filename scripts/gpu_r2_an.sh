#!/bin/bash
# Round 2, GPU call AN: edge-list grids sized to whole waves of the 148 SMs (SX_EDGE_BALANCE=1, default) against blocks of exactly ROWS rows.
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_edgelist_gpu.py tests/test_spmm_gpu.py -x -q -m gpu -p no:cacheprovider ) > gpurun_out/r2an_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2an_pytest.log
for b in 0 1; do
  SX_EDGE_BALANCE=$b timeout 600 python bench.py --configs pcrystk02_n8,pcrystk02_n16,pcrystk02_n32,pcrystk02_n64 --no-cpu-baseline --batch 0 > gpurun_out/r2an_b$b.json 2> gpurun_out/r2an_b$b.err
  python - $b <<'PY'
import json,sys
b=sys.argv[1]
try:
    d=json.load(open(f'gpurun_out/r2an_b{b}.json'))
    print('balance',b,'headline us', round(d['ms_per_step']*1e3,3), 'e2e', round(d['e2e']['ms_per_step']*1e3,1))
    for k,x in d['configs'].items(): print('   ',k, x['ms'], x['frac'], x['bit_exact'], x['kernel'][:75])
except Exception as e: print(b,'failed',e)
PY
done
