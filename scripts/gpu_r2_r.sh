#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_edgelist_gpu.py -x -q -p no:cacheprovider ) > gpurun_out/r2r_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2r_pytest.log
for hf in -1 0; do python bench.py --configs none --no-cpu-baseline --host-fused $hf > gpurun_out/r2r_bench_hf$hf.json 2>/dev/null; python -c "import json; d=json.load(open('gpurun_out/r2r_bench_hf$hf.json')); print('host-fused $hf: kernel us', d['ms_per_step']*1e3, 'e2e us', d['e2e']['ms_per_step']*1e3, d['e2e']['path'][:50])"; done
python bench.py --configs none --no-cpu-baseline --workload pcrystk02 --ncols 8 > gpurun_out/r2r_pcr8.json 2>/dev/null; python -c "import json; d=json.load(open('gpurun_out/r2r_pcr8.json')); print('pcr N=8: kernel us', d['ms_per_step']*1e3, 'e2e us', d['e2e']['ms_per_step']*1e3, d['e2e']['path'][:50])"
