#!/bin/bash
# Round 2, GPU call BD: 16-lane shapes of the edge-list kernel (256-byte rows) as ONE launch against two passes of 8-lane launches
# (SX_OPT_PANEL_COLS = 128 bytes of columns), after the row-aligned walk made the 8-lane kernel faster.
mkdir -p gpurun_out
for pc in 0 32; do
  timeout 600 python bench.py --configs pcrystk02_n64 --no-cpu-baseline --no-pipelined-e2e --batch 0 --panel-cols $pc > gpurun_out/r2bd_f32_pc$pc.json 2>/dev/null
  python -c "import json; d=json.load(open('gpurun_out/r2bd_f32_pc$pc.json')); x=d['configs']['pcrystk02_n64']; print('pcrystk02 N=64 f32 panel_cols=$pc:', x['ms']*1e3, 'us', x['bit_exact'], x['kernel'][:50])"
done
for pc in 0 16; do
  timeout 600 python bench.py --workload nasa4704 --ncols 32 --dtype f64 --configs none --no-cpu-baseline --no-pipelined-e2e --batch 0 --panel-cols $pc > gpurun_out/r2bd_f64_pc$pc.json 2>/dev/null
  python -c "import json; d=json.load(open('gpurun_out/r2bd_f64_pc$pc.json')); print('nasa4704 N=32 f64 panel_cols=$pc:', d['ms_per_step']*1e3, 'us', d['parity']['bit_exact_all_ranks'], d['roofline']['kernel'][:50])"
  timeout 600 python bench.py --workload pcrystk02 --ncols 32 --dtype f64 --configs none --no-cpu-baseline --no-pipelined-e2e --batch 0 --panel-cols $pc > gpurun_out/r2bd_p64_pc$pc.json 2>/dev/null
  python -c "import json; d=json.load(open('gpurun_out/r2bd_p64_pc$pc.json')); print('pcrystk02 N=32 f64 panel_cols=$pc:', d['ms_per_step']*1e3, 'us', d['parity']['bit_exact_all_ranks'], d['roofline']['kernel'][:50])"
done
