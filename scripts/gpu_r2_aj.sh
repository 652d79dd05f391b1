#!/bin/bash
# Round 2, GPU call AJ: the round's evidence on one GPU -- full GPU suite, compute-sanitizer over this session's kernels,
# ncu launch list of the default bench command, ncu --set full of the batched edge-list launch and of the one-kernel host
# call, the bench lines of both arms.
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider ) > gpurun_out/r2aj_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2aj_pytest.log
SEL='(column_pipeline and (997 or 333 or 1500)) or (batched and (997 or 70-64)) or (staged_B and (997 or 333)) or capture'
for tool in memcheck racecheck synccheck; do
  extra=""; [ $tool != memcheck ] && extra="--num-cuda-barriers 65536"
  timeout 900 compute-sanitizer --tool $tool $extra --error-exitcode 9 python -m pytest tests/test_edgelist_gpu.py -q -p no:cacheprovider -k "$SEL" > gpurun_out/r2aj_sanitizer_$tool.log 2>&1
  echo "compute-sanitizer $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/r2aj_sanitizer_$tool.log | tail -3
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2aj_launches_bench.csv python bench.py --steps 20 --warmup 5 --configs none --no-cpu-baseline --min-region-ms 0.05 > gpurun_out/r2aj_launches_bench.json 2> /dev/null; echo "launch list rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:spmm_edgelist_kernel -s 1 -c 1 -f -o gpurun_out/r2aj_full_batched python scripts/batch_probe.py > gpurun_out/r2aj_full_batched.log 2>&1; echo "ncu full batched rc=$?"
MTX=$(python -c "from sextans_b200 import workloads as w; print(w.suitesparse_path('nasa4704'))")
timeout 600 ncu --set full --import-source on --clock-control none -k regex:spmm_edgelist_host_kernel -s 40 -c 1 -f -o gpurun_out/r2aj_full_hostcall scripts/micro/e2e_c $MTX 16 > gpurun_out/r2aj_full_hostcall.log 2>&1; echo "ncu full host call rc=$?"; tail -2 gpurun_out/r2aj_full_hostcall.log
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2aj_bench_ref.json 2> gpurun_out/r2aj_bench_ref.err; echo "reference arm rc=$?"
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2aj_bench.json 2> gpurun_out/r2aj_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2aj_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2aj_bench.json')); r=json.load(open('gpurun_out/r2aj_bench_ref.json'))
print('headline us', d['ms_per_step']*1e3, 'frac', d['roofline']['frac'], 'e2e us', d['e2e']['ms_per_step']*1e3, 'e2e GF', d['e2e']['value'], 'ref GF', r['value'], 'ratio', d['e2e']['value']/r['value'])
print(d['run']['k_step_graphs']); print('batched', d['batched'])
for k,v in d['configs'].items(): print(k, v.get('ms'), v.get('frac'), v.get('parity'), str(v.get('kernel'))[:60], v.get('error',''))
PY
ls -la gpurun_out/r2aj_*
