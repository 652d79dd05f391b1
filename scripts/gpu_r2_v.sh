#!/bin/bash
# Round 2, GPU call V: column pipeline of the host-facing call (SX_OPT_HOST_GROUPS), batched call, parity.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -x -q -m gpu -p no:cacheprovider --deselect tests/test_baseline_configs_gpu.py ) > gpurun_out/r2v_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2v_pytest.log
for g in 1 2 4 8; do
  timeout 600 python bench.py --configs none --no-cpu-baseline --host-groups $g --batch $((g==1?20:0)) > gpurun_out/r2v_hg$g.json 2> gpurun_out/r2v_hg$g.err; echo "hg=$g rc=$?"
  python - gpurun_out/r2v_hg$g.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print('  headline us', round(d['ms_per_step']*1e3,3), 'e2e us', round(d['e2e']['ms_per_step']*1e3,2), 'checksum', d['checksum_C'], 'batched', d.get('batched'))
PY
done
timeout 600 python bench.py --workload pcrystk02 --ncols 16 --configs none --no-cpu-baseline --host-groups 1 --batch 0 > gpurun_out/r2v_pcr_hg1.json 2>/dev/null
timeout 600 python bench.py --workload pcrystk02 --ncols 16 --configs none --no-cpu-baseline --host-groups 4 --batch 20 > gpurun_out/r2v_pcr_hg4.json 2>/dev/null
python - <<'PY'
import json
for g in (1,4):
    d=json.load(open(f'gpurun_out/r2v_pcr_hg{g}.json'))
    print('pcrystk02 N=16 hg',g,'e2e us', round(d['e2e']['ms_per_step']*1e3,2), d['e2e']['path'][:40], 'batched', d.get('batched'))
PY
