#!/bin/bash
# Round 2, GPU call AU: the persistent batch kernel (spmm_edgelist_batch_kernel) against grid.y = nb launches of the single kernel.
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_edgelist_gpu.py -x -q -m gpu -p no:cacheprovider -k "batched" ) > gpurun_out/r2au_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2au_pytest.log
for p in 0 1; do for wl in "nasa4704 16" "pcrystk02 16" "pcrystk02 32"; do set -- $wl
  SX_BATCH_PERSISTENT=$p timeout 600 python bench.py --workload $1 --ncols $2 --configs none --no-cpu-baseline --no-pipelined-e2e --batch 20 > gpurun_out/r2au_tmp.json 2> gpurun_out/r2au_tmp.err
  python - $p $1 $2 <<'PY'
import json,sys
try:
    d=json.load(open('gpurun_out/r2au_tmp.json')); b=d['batched']
    print('persistent',sys.argv[1], sys.argv[2], 'N='+sys.argv[3], 'single us', round(d['ms_per_step']*1e3,3), 'batched us per SpMM', b.get('ms_per_spmm',0)*1e3, 'frac', b.get('frac'), b.get('kernel'), b.get('bit_exact_every_triple'), b.get('error',''))
except Exception as e: print('failed', e)
PY
done; done
for nb in 4 8 40; do
  timeout 600 python bench.py --configs none --no-cpu-baseline --no-pipelined-e2e --batch $nb > gpurun_out/r2au_tmp.json 2>/dev/null
  python -c "import json; b=json.load(open('gpurun_out/r2au_tmp.json'))['batched']; print('nasa4704 nb=$nb', b['ms_per_spmm']*1e3, 'us per SpMM, frac', b['frac'], b['bit_exact_every_triple'])"
done
