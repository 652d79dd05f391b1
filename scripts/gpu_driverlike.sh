#!/bin/bash
# what the driver does at round end: GPU tests, smoke, reference arm, own arm
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -x -q -m gpu ) > gpurun_out/dl_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/dl_pytest.log
( time python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/dl_smoke.log 2>&1; echo "smoke rc=$?"; tail -6 gpurun_out/dl_smoke.log
( time python bench.py --impl reference ) > gpurun_out/dl_bench_ref.log 2>&1; echo "ref rc=$?"; grep -v "^$" gpurun_out/dl_bench_ref.log | tail -5 | cut -c1-1500
( time python bench.py ) > gpurun_out/dl_bench.log 2>&1; echo "bench rc=$?"; tail -5 gpurun_out/dl_bench.log | cut -c1-3000
