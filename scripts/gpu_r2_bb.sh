#!/bin/bash
# Round 2, GPU call BB: compute-sanitizer over the edge-list kernels after the row-aligned streams (memcheck, racecheck, synccheck).
mkdir -p gpurun_out
SEL='small_golden or every_kernel_variant or 1000-1000-8 or 70-64-4 or 333-777-1 or canned or dependent_chain or 999-1200-32 or (column_pipeline and (997 or 333)) or (batched and 70-64) or (staged_B and 333) or in_flight'
for tool in memcheck racecheck synccheck; do
  extra=""; [ $tool != memcheck ] && extra="--num-cuda-barriers 65536"
  timeout 900 compute-sanitizer --tool $tool $extra --error-exitcode 9 python -m pytest tests/test_spmm_gpu.py tests/test_edgelist_gpu.py -q -p no:cacheprovider -k "$SEL" > gpurun_out/r2bb_sanitizer_$tool.log 2>&1
  echo "compute-sanitizer $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/r2bb_sanitizer_$tool.log | tail -2
done
grep "Race reported between" gpurun_out/r2bb_sanitizer_racecheck.log | sed 's/+0x[0-9a-f]*//g; s/=========//; s/^[. ]*//' | sort | uniq -c | sort -rn | cut -c1-200
