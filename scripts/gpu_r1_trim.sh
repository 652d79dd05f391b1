#!/bin/bash
# after trimming the staged kernel's batch loop (shift/mask instead of divisions, one
# IMAD.WIDE per gather address): parity of every variant, then the C5-like probe
mkdir -p gpurun_out
( time timeout 200 python -m pytest tests/test_spmm_gpu.py tests/test_windows_gpu.py -x -q -k "not device_resident" -p no:cacheprovider ) > gpurun_out/r1_trim_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r1_trim_tests.log; tail -6 gpurun_out/r1_trim_tests.log
PROBE_W=0 PROBE_REPS=10 timeout 100 python scripts/probe_windows.py > gpurun_out/r1_trim_probe_default.log 2>&1; cat gpurun_out/r1_trim_probe_default.log
SX_STAGE_KB=56 PROBE_W=0,262144 PROBE_REPS=10 timeout 100 python scripts/probe_windows.py > gpurun_out/r1_trim_probe_stage56.log 2>&1; cat gpurun_out/r1_trim_probe_stage56.log
