#!/bin/bash
set -u
O=gpurun_out/s5d; mkdir -p $O
for v in new oldred; do
  if [ $v = oldred ]; then export GCB_LIB=$PWD/gochem_b200/lib/libgochem_b200_oldred.so; fi
  {
  echo "== lovo 50kx5k";   timeout 120 python tools/tune.py 5000 50000 lovo 0,0,0,-1
  echo "== lovo 100kx10k"; timeout 120 python tools/tune.py 10000 100000 lovo 0,0,0,-1
  echo "== fit 100kx1500"; timeout 120 python tools/tune.py 1500 100000 fit 0,0,0,-1
  echo "== fit 100kx3000"; timeout 120 python tools/tune.py 3000 100000 fit 0,0,0,-1
  echo "== fit 50kx5k";    timeout 120 python tools/tune.py 5000 50000 fit 0,0,0,-1
  echo "== fit 50kx20k";    timeout 120 python tools/tune.py 20000 50000 fit 0,0,0,-1
  echo "== wb 100kx10k";   timeout 120 python tools/tune.py 10000 100000 wb 0,0,0,-1
  echo "== wb 100kx2000";  timeout 120 python tools/tune.py 2000 100000 wb 0,0,0,-1
  echo "== lovo 100kx2000"; timeout 120 python tools/tune.py 2000 100000 lovo 0,0,0,-1
  echo "== wb 100kx1500";  timeout 120 python tools/tune.py 1500 100000 wb 0,0,0,-1
  } > $O/tune_$v.log 2>&1
done
