#!/bin/bash
# Round 2, GPU call AE: the bench line with every K-step window chained into one graph (steady-state chain) next to the
# K-step-graph figure; the new capture test.
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_edgelist_gpu.py -x -q -m gpu -p no:cacheprovider -k "capture or canned" ) > gpurun_out/r2ae_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2ae_pytest.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2ae_bench.json 2> gpurun_out/r2ae_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2ae_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2ae_bench.json'))
print('headline us', d['ms_per_step']*1e3, 'frac', d['roofline']['frac'], 'e2e us', d['e2e']['ms_per_step']*1e3)
print(d['run']['timed']); print(d['run']['launch']); print(d['run']['k_step_graphs'])
print('batched', d['batched'])
for k,v in d['configs'].items(): print(k, v.get('ms'), v.get('frac'), v.get('parity'), str(v.get('kernel'))[:60], v.get('error',''))
print('launches', d['gpu_launches'], d['gpu_launches_detail'])
PY
