#!/bin/bash
# Build a tuning variant of the library: tools/build_variant.sh <name> [nvcc -D flags...]  -> build/variants/lib_<name>.so
# (selected at run time with HMVIT_LIB=...; build/ is git-ignored but travels to the GPU box)
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
name="$1"; shift
mkdir -p "$ROOT/build/variants"
cd "$ROOT/hm-vit_b200"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --shared -Xcompiler -fPIC -Xptxas -v "$@" \
  -o "$ROOT/build/variants/lib_$name.so" csrc/api.cu > "$ROOT/build/variants/$name.log" 2>&1 || { grep -i error "$ROOT/build/variants/$name.log"; exit 1; }
echo "built $name"
