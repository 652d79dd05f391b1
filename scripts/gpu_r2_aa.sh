#!/bin/bash
# Round 2, GPU call AA: staged kernel on C5 -- register/occupancy variants (build-time knobs) and work-item sizes.
mkdir -p gpurun_out
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    for k,v in d['configs'].items(): print(sys.argv[2], k, v.get('ms'), v.get('frac'), v.get('parity'), str(v.get('kernel'))[:70], v.get('error',''))
except Exception as e: print(sys.argv[2], 'failed', e)
PY
}
run() { tag=$1; shift; timeout 600 python bench.py --configs powerlaw_c5 --no-cpu-baseline --batch 0 "$@" > gpurun_out/r2aa_$tag.json 2> gpurun_out/r2aa_$tag.err; show gpurun_out/r2aa_$tag.json $tag; }
run default
SX_LIBRARY_PATH=$PWD/sextans_b200/variants/libsextans_b200_minb4.so run minb4
SX_LIBRARY_PATH=$PWD/sextans_b200/variants/libsextans_b200_umax16b2.so run umax16b2
run item256 --item-nnz 256
run item1024 --item-nnz 1024
run nopf --prefetch 0
SX_LIBRARY_PATH=$PWD/sextans_b200/variants/libsextans_b200_minb4.so run minb4_item1024 --item-nnz 1024
