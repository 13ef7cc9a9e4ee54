#!/bin/bash
set -u
O=gpurun_out/s5x; mkdir -p $O
for na in 700 1000 1500; do for mode in wb lovo; do
  echo "== $mode $na"; timeout 120 python tools/tune.py $na 100000 $mode 0,0,0,-1 1,0,1,1 1,0,2,1 1,0,3,1 2,0,2,1 2,0,3,1 | grep -v generic
done; done > $O/tune.log 2>&1
