#!/bin/bash
# final check of round 1: the driver's own GPU test command, smoke, then the C5 bench line
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests -x -q -m gpu -p no:cacheprovider ) > gpurun_out/r1_final2_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r1_final2_pytest.log; tail -6 gpurun_out/r1_final2_pytest.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1_final2_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/r1_final2_smoke.log; tail -4 gpurun_out/r1_final2_smoke.log
timeout 150 python bench.py --workload powerlaw --steps 10 --no-cpu-baseline > gpurun_out/r1_final2_bench_powerlaw.json 2> gpurun_out/r1_final2_bench_powerlaw.err; echo "bench exit $?"; cut -c1-1200 gpurun_out/r1_final2_bench_powerlaw.json
