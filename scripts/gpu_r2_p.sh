#!/bin/bash
# Round 2, GPU call P: memcheck after the fix of the staged kernel's idle-lane gathers; the whole GPU suite; the bench.
mkdir -p gpurun_out
SEL='small_golden or every_kernel_variant or 1000-1000-8 or 70-64-4 or 333-777-1 or canned or dependent_chain or 999-1200-32 or l2_prefetch or config2'
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_spmm_gpu.py tests/test_edgelist_gpu.py tests/test_optin_kernels_gpu.py tests/test_windows_gpu.py tests/test_images_gpu.py -q -p no:cacheprovider -k "$SEL or windows or images" > gpurun_out/r2p_sanitizer_memcheck.log 2>&1
echo "compute-sanitizer memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2p_sanitizer_memcheck.log | tail -3
( time timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider ) > gpurun_out/r2p_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2p_pytest.log
( time python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; echo "bench rc=$?"; tail -4 gpurun_out/r2p_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2p_bench.json'))
print('headline us', d['ms_per_step']*1e3, 'frac', d['roofline']['frac'], 'traffic', d['roofline']['traffic'], 'e2e us', d['e2e']['ms_per_step']*1e3, d['roofline']['kernel'])
for k,v in d['configs'].items(): print(k, v['ms'], v['frac'], v['traffic_ratio'], v['parity'], v['kernel'][:70])
print(d['cpu_baseline'])
PY
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2p_bench_ref.json 2>/dev/null; cut -c1-300 gpurun_out/r2p_bench_ref.json
python bench.py --configs powerlaw_blocked --no-cpu-baseline > gpurun_out/r2p_bench_blocked.json 2> gpurun_out/r2p_bench_blocked.err; python -c "import json; d=json.load(open('gpurun_out/r2p_bench_blocked.json')); print(d['configs'])"
python bench.py --workload fem --configs none --no-cpu-baseline --steps 10 > gpurun_out/r2p_fem2000.json 2>/dev/null; python -c "import json; d=json.load(open('gpurun_out/r2p_fem2000.json')); print('fem band 2000:', d['ms_per_step'], d['roofline']['kernel'][:60])"
python bench.py --workload fem --band 100 --configs none --no-cpu-baseline --steps 10 > gpurun_out/r2p_fem100.json 2>/dev/null; python -c "import json; d=json.load(open('gpurun_out/r2p_fem100.json')); print('fem band 100:', d['ms_per_step'], d['roofline']['kernel'][:60])"
