#!/bin/bash
set -u
O=gpurun_out/s5z3; mkdir -p $O
for g in default 32 64; do
  if [ $g = default ]; then unset GCB_L2_FETCH; else export GCB_L2_FETCH=$g; fi
  timeout 120 python tools/subset_fit_probe.py 100000 20000 300 1000 3000 > $O/probe_$g.json 2>> $O/err.log
  timeout 120 python tools/tune.py 10000 100000 fit 0,0,0,-1 2>> $O/err.log | grep cfg | grep -v generic > $O/fit_$g.json
done
