#!/bin/bash
# Tiny and degenerate inputs through the drop-in CLI (both modes); every output must decode to its input.
cli=megalania_b200/_build/megalania
for content in "hello hello" "a" "ab" "abcdefgh" "aaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaa" "abababababababababababababababab"; do
  printf '%s' "$content" > /tmp/edge.bin
  for mode in "--rounds 3 --round-ms 20" "--iters 200 --epochs 1 --steps 2 --chains 64"; do
    timeout 60 $cli $mode /tmp/edge.bin > /tmp/edge.lzma 2> /tmp/edge.err; rc=$?
    if [ $rc -ne 0 ]; then echo "FAIL rc=$rc [$content] [$mode]: $(tail -1 /tmp/edge.err)"; continue; fi
    if xz --format=lzma -dc /tmp/edge.lzma | cmp -s - /tmp/edge.bin; then echo "ok   [$content] [$mode] -> $(wc -c < /tmp/edge.lzma) bytes"; else echo "FAIL roundtrip [$content] [$mode]"; fi
  done
done
: > /tmp/empty.bin; timeout 20 $cli /tmp/empty.bin > /tmp/e.lzma 2>/tmp/e.err; echo "empty file: rc=$? out=$(wc -c < /tmp/e.lzma) $(tail -1 /tmp/e.err)"
