#!/bin/bash
# Round 2, GPU call I: full -m gpu suite after the clean-up, fused host path (e2e), bench.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -x -q -m gpu -p no:cacheprovider ) > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2i_pytest.log
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2i_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2i_bench.json'))
print('headline us', d['ms_per_step']*1e3, 'frac', d['roofline']['frac'], 'e2e us', d['e2e']['ms_per_step']*1e3, d['e2e']['path'][:60], d['run']['timed'], d['roofline']['kernel'])
for k,v in d['configs'].items(): print(k, v['ms'], v['frac'], v['parity'], v['kernel'][:70])
PY
for hf in 0; do python bench.py --configs none --no-cpu-baseline --host-fused $hf > gpurun_out/r2i_bench_hf$hf.json 2>/dev/null; python -c "import json; d=json.load(open('gpurun_out/r2i_bench_hf$hf.json')); print('host-fused $hf: e2e us', d['e2e']['ms_per_step']*1e3, d['e2e']['path'][:50])"; done
for wl in "pcrystk02 --ncols 16" "pcrystk02 --ncols 64"; do python bench.py --configs none --no-cpu-baseline --workload $wl > gpurun_out/r2i_tmp.json 2>/dev/null; python -c "import json; d=json.load(open('gpurun_out/r2i_tmp.json')); print('$wl: kernel us', d['ms_per_step']*1e3, 'e2e us', d['e2e']['ms_per_step']*1e3, d['e2e']['path'][:40])"; done
