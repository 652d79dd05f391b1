#!/bin/bash
# Round 2, GPU call AR: two host-facing calls in flight (sx_spmm_enqueue_* on two contexts): test and the e2e.pipelined figure.
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_edgelist_gpu.py -x -q -m gpu -p no:cacheprovider -k "in_flight or column_pipeline" ) > gpurun_out/r2ar_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2ar_pytest.log
timeout 600 python bench.py --configs none --no-cpu-baseline --batch 0 > gpurun_out/r2ar_bench.json 2> gpurun_out/r2ar_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2ar_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2ar_bench.json'))
print('e2e blocking us', d['e2e']['ms_per_step']*1e3, 'GF', d['e2e']['value']); print('pipelined', d['e2e']['pipelined'])
PY
