#!/bin/bash
# A/B of library variants where the 3x3 solve matters: bash tools/ab_solver.sh <out-file> <variant suffixes ...>  ("" = product)
out=$1; shift
: > $out
for job in "5000 50000 lovo" "10000 100000 lovo" "10000 100000 fit" "10000 100000 wb" "1500 100000 fit" "2000 100000 lovo" "1000 100000 wb" "640 100000 lovo"; do
  for v in "$@"; do
    lib=gochem_b200/lib/libgochem_b200${v:+_$v}.so
    echo "=== variant '$v' job $job" >> $out
    GCB_LIB=$PWD/$lib timeout 300 python tools/tune.py $job 0,0,0,-1 2>&1 | grep -v "generic\|plan" >> $out
  done
done
for v in "$@"; do
  lib=gochem_b200/lib/libgochem_b200${v:+_$v}.so
  echo "=== variant '$v' small structures (tools/small_sweep.py)" >> $out
  GCB_LIB=$PWD/$lib timeout 300 python tools/small_sweep.py 2>&1 | tail -12 >> $out
done
