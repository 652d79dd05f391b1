#!/bin/bash
# Round 2, GPU call AC: narrow-N cases of the bench (N=4 fp32 failed in call AB: why), full GPU suite incl. the full-size configs.
mkdir -p gpurun_out
for n in 4 2 1; do
  ( time timeout 300 python bench.py --workload nasa4704 --ncols $n --dtype f32 --configs none --no-cpu-baseline --batch 4 --no-flush ) > gpurun_out/r2ac_n$n.json 2> gpurun_out/r2ac_n$n.err; echo "N=$n noflush rc=$?"; tail -3 gpurun_out/r2ac_n$n.err; head -c 300 gpurun_out/r2ac_n$n.json; echo
done
( time timeout 600 python bench.py --workload nasa4704 --ncols 4 --dtype f32 --configs none --no-cpu-baseline --batch 0 ) > gpurun_out/r2ac_n4_cold.json 2> gpurun_out/r2ac_n4_cold.err; echo "N=4 cold rc=$?"; tail -5 gpurun_out/r2ac_n4_cold.err; head -c 300 gpurun_out/r2ac_n4_cold.json; echo
( time timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider ) > gpurun_out/r2ac_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2ac_pytest.log
