#!/bin/bash
# Round 2, GPU call AB: programmatic dependent launch on/off for the edge-list kernel over lane-group widths and both
# SuiteSparse matrices (the auto rule of SX_OPT_PDL), cold-rotation timing as in the bench line.
mkdir -p gpurun_out
: > gpurun_out/r2ab_pdl.txt
for wl in nasa4704 pcrystk02; do for dt in f32 f64; do for n in 4 8 16 32; do for pdl in 0 1; do
  timeout 300 python bench.py --workload $wl --ncols $n --dtype $dt --pdl $pdl --configs none --no-cpu-baseline --batch 0 --min-region-ms 20 > gpurun_out/r2ab_tmp.json 2>/dev/null
  python - $wl $dt $n $pdl <<'PY' | tee -a gpurun_out/r2ab_pdl.txt
import json,sys
try:
    d=json.load(open('gpurun_out/r2ab_tmp.json'))
    print(*sys.argv[1:], 'us', round(d['ms_per_step']*1e3,3), d['roofline']['kernel'][:60], 'bit_exact', d['parity']['bit_exact_all_ranks'])
except Exception as e: print(*sys.argv[1:], 'failed', e)
PY
done; done; done; done
