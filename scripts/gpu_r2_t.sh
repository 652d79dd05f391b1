#!/bin/bash
# Round 2, GPU call T: super-rows in the edge-list kernel -- parity, then the bench with all configs.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -x -q -m gpu -p no:cacheprovider --deselect tests/test_baseline_configs_gpu.py ) > gpurun_out/r2t_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2t_pytest.log
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2t_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2t_bench.json'))
print('headline us', d['ms_per_step']*1e3, 'frac', d['roofline']['frac'], 'e2e us', d['e2e']['ms_per_step']*1e3, d['roofline']['kernel'])
for k,v in d['configs'].items(): print(k, v['ms'], v['frac'], v['parity'], v['kernel'][:70])
PY
python bench.py --configs none --no-cpu-baseline --pdl 0 > gpurun_out/r2t_nopdl.json 2>/dev/null; python -c "import json; d=json.load(open('gpurun_out/r2t_nopdl.json')); print('nopdl headline us', d['ms_per_step']*1e3)"
