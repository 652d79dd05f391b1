#!/bin/bash
# Round 2, GPU call AT: memcheck / racecheck / synccheck over the two-calls-in-flight test, then the whole GPU suite once more.
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_edgelist_gpu.py -q -p no:cacheprovider -k "in_flight" > gpurun_out/r2at_sanitizer_$tool.log 2>&1
  echo "compute-sanitizer $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/r2at_sanitizer_$tool.log | tail -2
done
( time timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider ) > gpurun_out/r2at_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2at_pytest.log
