#!/bin/bash
# Drop-in CLI, cooperative mode, against xz -9e on the same synthetic inputs:
#   tools/cli_probe.sh <kind> <size> <rounds> <round-ms> [group]
set -e
kind=$1; size=$2; rounds=$3; ms=$4; group=${5:-8}
f=/tmp/mg_$kind$size
python tools/corpus.py $kind $size -o $f
xzb=$(xz -9e --format=lzma -c $f | wc -c)
s=$(date +%s%N)
megalania_b200/_build/megalania --chains 4736 --rounds $rounds --round-ms $ms --group $group $f > $f.lzma 2> $f.log
e=$(date +%s%N)
xz --format=lzma -dc $f.lzma | cmp - $f
echo "$kind $size: megalania-b200 $(wc -c < $f.lzma) bytes in $(( (e - s) / 1000000 )) ms ($rounds rounds x $ms ms, group $group); xz -9e $xzb bytes; round-trips through xz -d"
tail -1 $f.log
