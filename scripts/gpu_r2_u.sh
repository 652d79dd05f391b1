#!/bin/bash
# Round 2, GPU call U: after the revert of super-rows -- parity, then N passes (SX_OPT_PANEL_COLS) on C4 / C5.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -x -q -m gpu -p no:cacheprovider --deselect tests/test_baseline_configs_gpu.py ) > gpurun_out/r2u_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2u_pytest.log
show() { python - "$1" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1],'headline us', round(d['ms_per_step']*1e3,3))
for k,v in d['configs'].items(): print('  ',k, v['ms'], v['frac'], v['parity'], v['kernel'][:70])
PY
}
for pc in 0 64 32 16; do
  timeout 600 python bench.py --configs uniform_c4 --no-cpu-baseline --panel-cols $pc > gpurun_out/r2u_c4_pc$pc.json 2> gpurun_out/r2u_c4_pc$pc.err; echo "c4 pc=$pc rc=$?"; show gpurun_out/r2u_c4_pc$pc.json
done
timeout 600 python bench.py --configs powerlaw_c5 --no-cpu-baseline --panel-cols 8 > gpurun_out/r2u_c5_pc8.json 2> gpurun_out/r2u_c5_pc8.err; echo "c5 pc=8 rc=$?"; show gpurun_out/r2u_c5_pc8.json
timeout 600 python bench.py --configs pcrystk02_n64 --no-cpu-baseline --panel-cols 32 > gpurun_out/r2u_pcr64_pc32.json 2> gpurun_out/r2u_pcr64_pc32.err; echo "pcr64 pc=32 rc=$?"; show gpurun_out/r2u_pcr64_pc32.json
