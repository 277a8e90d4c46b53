#!/bin/bash
# Are two builds of the library the same machine code, kernel by kernel?  (Used when experiment switches are removed
# from the source after a GPU run: the product build must equal the build the GPU tests ran on.)
#   tools/sass_same.sh megalania_b200/_build/libmegalania_cuda.so megalania_b200/_build/variants/lib_v0.so
a=$1; b=$2; rc=0
for k in $(cuobjdump -sass "$a" | sed -n 's/.*Function : \(.*\)$/\1/p'); do
  da=$(cuobjdump -sass "$a" | awk -v k="$k" '/Function : /{f=($NF==k)} f' | grep -v '^\s*$' | md5sum)
  db=$(cuobjdump -sass "$b" | awk -v k="$k" '/Function : /{f=($NF==k)} f' | grep -v '^\s*$' | md5sum)
  if [ "$da" = "$db" ]; then echo "same    $k"; else echo "DIFFERS $k"; rc=1; fi
done
exit $rc
