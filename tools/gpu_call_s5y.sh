#!/bin/bash
set -u
O=gpurun_out/s5y; mkdir -p $O
export GCB_LIB=$PWD/gochem_b200/lib/libgochem_b200_w4.so
for na in 500 700 1000 1500; do for mode in wb lovo; do
  echo "== $mode $na"; timeout 120 python tools/tune.py $na 100000 $mode 1,0,2,1 | grep -v generic
done; done > $O/tune_w4.log 2>&1
