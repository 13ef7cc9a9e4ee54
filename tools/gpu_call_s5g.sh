#!/bin/bash
set -u
O=gpurun_out/s5g; mkdir -p $O
for v in base ch2 ch10 tiny; do
  if [ $v = base ]; then unset GCB_LIB; else export GCB_LIB=$PWD/gochem_b200/lib/libgochem_b200_$v.so; fi
  {
  echo "== lovo 50kx5k";   timeout 120 python tools/tune.py 5000 50000 lovo 0,0,0,-1
  echo "== lovo 100kx10k"; timeout 120 python tools/tune.py 10000 100000 lovo 0,0,0,-1
  echo "== lovo 100kx2000"; timeout 120 python tools/tune.py 2000 100000 lovo 0,0,0,-1
  } > $O/tune_$v.log 2>&1
done
