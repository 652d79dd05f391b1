#!/bin/bash
# Round 2, GPU call F: phase timeline of the edge-list kernel; the rewritten bench.py at N=1 with every config.
mkdir -p gpurun_out
for pdl in -1 0; do
  SX_LIBRARY_PATH=$PWD/sextans_b200/variants/libsextans_b200_trace.so python scripts/edge_trace.py --pdl $pdl > gpurun_out/r2f_trace_nasa_pdl$pdl.txt 2>&1; tail -28 gpurun_out/r2f_trace_nasa_pdl$pdl.txt
done
SX_LIBRARY_PATH=$PWD/sextans_b200/variants/libsextans_b200_trace.so python scripts/edge_trace.py --workload pcrystk02 --copies 30 > gpurun_out/r2f_trace_pcr.txt 2>&1; tail -8 gpurun_out/r2f_trace_pcr.txt
( time python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"; tail -5 gpurun_out/r2f_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2f_bench.json'))
print('headline us', d['ms_per_step']*1e3, 'frac', d['roofline']['frac'], 'e2e us', d['e2e']['ms_per_step']*1e3, d['run']['timed'])
print('parity', d['parity'])
for k,v in d['configs'].items(): print(k, v)
print('cpu', d.get('cpu_baseline'))
PY
( time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2f_bench_ref.json 2> gpurun_out/r2f_bench_ref.err; echo "ref rc=$?"; cut -c1-600 gpurun_out/r2f_bench_ref.json
