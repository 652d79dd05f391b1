#!/bin/bash
# first GPU pass: parity tests, smoke, bench lines, launch list, one full ncu capture
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/smoke.log
python bench.py > gpurun_out/bench_nasa4704.json 2> gpurun_out/bench_nasa4704.err; echo "bench rc=$?"; cat gpurun_out/bench_nasa4704.json; tail -3 gpurun_out/bench_nasa4704.err
python bench.py --impl reference > gpurun_out/bench_ref_nasa4704.json 2>&1; cat gpurun_out/bench_ref_nasa4704.json
for wlk in pcrystk02 uniform powerlaw; do
  python bench.py --workload $wlk --steps 20 > gpurun_out/bench_$wlk.json 2> gpurun_out/bench_$wlk.err; echo "bench $wlk rc=$?"; cat gpurun_out/bench_$wlk.json; tail -2 gpurun_out/bench_$wlk.err
done
./sextans_b200/sextans /tmp/sextans_b200_fixtures/nasa4704.mtx 16 100 --json > gpurun_out/sextans_nasa4704.log 2>&1; tail -8 gpurun_out/sextans_nasa4704.log
./sextans_b200/sextans /tmp/sextans_b200_fixtures/nasa4704.mtx 16 100 --dtype f64 --json > gpurun_out/sextans_nasa4704_f64.log 2>&1; tail -4 gpurun_out/sextans_nasa4704_f64.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_nasa4704.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spmm_rows -s 3 -c 2 -o gpurun_out/prof_nasa4704 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spmm_ -s 3 -c 3 -o gpurun_out/prof_powerlaw python bench.py --workload powerlaw --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spmm_rows -s 3 -c 2 -o gpurun_out/prof_uniform python bench.py --workload uniform --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full3.log 2>&1
ls -la gpurun_out
