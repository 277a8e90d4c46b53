#!/bin/bash
# Kernel A/B runs on ONE box: builds the library with each -D set given and times the same short bench-like run.
#   tools/variants.sh "-DMG_PAIR_VARIANT=0" "-DMG_PAIR_VARIANT=1" ...        (run under gpurun)
set -e
cd "$(dirname "$0")/.."
mkdir -p megalania_b200/_build/variants gpurun_out
i=0
for defs in "$@"; do
  out=megalania_b200/_build/variants/lib_$i.so
  if [ ! -f "$out" ]; then
    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --shared -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -ldl $defs -o "$out" megalania_b200/csrc/mg_api.cu
  fi
  echo "== variant $i: $defs"
  MEGALANIA_CUDA_LIB=$out python tools/variant_run.py
  i=$((i+1))
done
