#!/bin/bash
# Round 2, GPU call X: the host-facing call as ONE kernel (spmm_edgelist_host_kernel): parity, timing from C, the bench line.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -x -q -m gpu -p no:cacheprovider --deselect tests/test_baseline_configs_gpu.py ) > gpurun_out/r2x_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2x_pytest.log
MTX=$(python -c "from sextans_b200 import workloads as w; print(w.suitesparse_path('nasa4704'))")
( for g in 1 2 4 8; do scripts/micro/e2e_c $MTX 16 $g 2; done
  scripts/micro/e2e_c $MTX 16 1 1
  scripts/micro/e2e_c $MTX 8 4 2; scripts/micro/e2e_c $MTX 8 1 1 ) 2>&1 | tee gpurun_out/r2x_e2e_c.txt
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_edgelist_gpu.py -x -q -p no:cacheprovider -k "column_pipeline and 2-float64 and (4704 or 1500)" > gpurun_out/r2x_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r2x_memcheck.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2x_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2x_bench.json'))
print('headline us', d['ms_per_step']*1e3, 'frac', d['roofline']['frac'], 'e2e us', d['e2e']['ms_per_step']*1e3, d['e2e']['path'][:60])
print('batched', d['batched'])
for k,v in d['configs'].items(): print(k, v['ms'], v['frac'], v['parity'], v['kernel'][:70])
print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
PY
