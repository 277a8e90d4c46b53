#!/bin/bash
# Kernel A/B runs on ONE box: builds the library with each -D set given and runs a script on it.
#   tools/variants.sh [-s script.py] "-DMG_WALK_VARIANT=0" "-DMG_WALK_VARIANT=1" ...        (run under gpurun)
set -e
cd "$(dirname "$0")/.."
script=tools/variant_run.py
if [ "$1" = "-s" ]; then script=$2; shift 2; fi
mkdir -p megalania_b200/_build/variants gpurun_out
i=0
for defs in "$@"; do
  out=megalania_b200/_build/variants/lib_$i.so
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --shared -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -ldl $defs -o "$out" megalania_b200/csrc/mg_api.cu
  echo "== variant $i: $defs"
  MEGALANIA_CUDA_LIB=$out python $script
  i=$((i+1))
done
