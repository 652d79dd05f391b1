#!/bin/bash
# Build an alternative libsextans_b200 with build-time knobs set, next to the default one:
#   scripts/build_variant.sh umax16 "-DSX_STAGED_UMAX=16 -DSX_STAGED_MINBLOCKS_F64=2"
# -> sextans_b200/variants/libsextans_b200_umax16.so   (git-ignored, travels with gpurun)
# Select it at run time with SX_LIBRARY_PATH=<that file> (sextans_b200/__init__.py).
# Run this HERE (CPU, ~1.5 min) before a gpurun call -- not on the GPU box.
set -e
name=$1; defs=$2
root=$(cd "$(dirname "$0")/.." && pwd)
src=$root/sextans_b200/csrc
out=$root/sextans_b200/variants
obj=$(mktemp -d)
mkdir -p "$out"
make -C "$src" sx_host.o sx_images.o > /dev/null
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -ccbin /usr/bin/g++ \
    $defs -Xptxas -v -c -o "$obj/sx_api.o" "$src/sx_api.cu" 2> "$out/ptxas_$name.log"
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$out/libsextans_b200_$name.so" \
    "$obj/sx_api.o" "$src/sx_host.o" "$src/sx_images.o" -cudart static
rm -rf "$obj"
echo "built $out/libsextans_b200_$name.so ($defs)"
