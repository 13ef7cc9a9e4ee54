#!/bin/bash
# A/B of library variants over the cluster kernel's main shapes: bash tools/ab_modes.sh <out-file> <variant suffixes ...>  ("" = product)
out=$1; shift
: > $out
for job in "5000 50000 lovo" "10000 100000 lovo" "10000 100000 fit" "10000 100000 wb" "1500 100000 fit" "20000 50000 fit" "2000 100000 lovo"; do
  for v in "$@"; do
    lib=gochem_b200/lib/libgochem_b200${v:+_$v}.so
    echo "=== variant '$v' job $job" >> $out
    GCB_LIB=$PWD/$lib timeout 300 python tools/tune.py $job 0,0,0,-1 2>&1 | grep -v "generic\|plan" >> $out
  done
done
