#!/bin/bash
# A/B of the per-atom-distance (LOVO) pass: 8 against 12 consumer warps (tuning hook resident = 1 | warps << 4) over frame sizes.
#   bash tools/ab_lovo.sh <out-file> [variant-suffix]
out=$1; v=$2
lib=gochem_b200/lib/libgochem_b200${v:+_$v}.so
: > $out
for shape in "5000 50000" "10000 100000" "20000 50000" "3000 100000" "2000 100000" "1500 100000" "1000 100000" "640 100000"; do
  echo "=== shape $shape (cfg = cluster,stages,depth,resident|warps<<4)" >> $out
  GCB_LIB=$PWD/$lib timeout 300 python tools/tune.py $shape lovo 0,0,0,129 0,0,0,193 0,0,0,-1 2>&1 | grep -v generic >> $out
done
