#!/bin/bash
# L2 prefetch of the next batch's B rows: parity, then C5-like and C4-like probes
mkdir -p gpurun_out
( time timeout 200 python -m pytest tests/test_spmm_gpu.py tests/test_windows_gpu.py -x -q -k "not device_resident" -p no:cacheprovider ) > gpurun_out/r1_pf_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r1_pf_tests.log; tail -6 gpurun_out/r1_pf_tests.log
PROBE_SET=0:0,0:1,262144:1 timeout 100 python scripts/probe_windows.py > gpurun_out/r1_pf_probe_powerlaw.log 2>&1; cat gpurun_out/r1_pf_probe_powerlaw.log
PROBE_KIND=uniform PROBE_SET=0:0,0:1 timeout 100 python scripts/probe_windows.py > gpurun_out/r1_pf_probe_uniform.log 2>&1; cat gpurun_out/r1_pf_probe_uniform.log
