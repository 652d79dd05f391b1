#!/bin/bash
# Last GPU call of round 1 (a few minutes of budget): the new GPU parity tests first, then
# the column-window probe.  Everything is written under gpurun_out/ as it goes.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/r1_last_gpu.txt 2>&1
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1_last_smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/r1_last_smoke.log
tail -4 gpurun_out/r1_last_smoke.log
timeout 110 python -m pytest tests/test_images_gpu.py tests/test_windows_gpu.py -q -k "not device_resident" -p no:cacheprovider > gpurun_out/r1_last_new_tests.log 2>&1
echo "new tests exit $?" >> gpurun_out/r1_last_new_tests.log
tail -3 gpurun_out/r1_last_new_tests.log
timeout 150 python scripts/probe_windows.py > gpurun_out/r1_last_probe_windows.log 2>&1
echo "probe exit $?" >> gpurun_out/r1_last_probe_windows.log
cat gpurun_out/r1_last_probe_windows.log
timeout 120 python -m pytest tests/test_windows_gpu.py -q -k "device_resident" -p no:cacheprovider > gpurun_out/r1_last_torch_test.log 2>&1
echo "torch test exit $?" >> gpurun_out/r1_last_torch_test.log
tail -3 gpurun_out/r1_last_torch_test.log
