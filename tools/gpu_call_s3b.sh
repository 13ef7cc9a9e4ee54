#!/bin/bash
# session-3 call B: LOVO pass after the masked/cubic-sqrt rework (8 and 12 consumer warps)
set -u
O=gpurun_out/s3b; mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 300 python tools/tune.py 5000 50000 lovo 2,0,1,1 2,0,2,1 3,0,1,1 3,0,2,1 4,0,2,1 2,2,2,0 2,3,3,0 > $O/tune_lovo_cw8.log 2>&1
GCB_LIB=$PWD/gochem_b200/lib/libgochem_b200_cw12.so timeout 300 python tools/tune.py 5000 50000 lovo 2,0,1,1 2,0,2,1 3,0,1,1 3,0,2,1 4,0,2,1 > $O/tune_lovo_cw12.log 2>&1
timeout 300 python tools/tune.py 10000 100000 lovo 3,0,1,1 4,0,1,1 4,0,2,1 5,0,2,1 > $O/tune_lovo10k_cw8.log 2>&1
timeout 300 python tools/tune.py 10000 100000 wb 4,0,1,1 4,0,2,1 5,0,2,1 > $O/tune_wb10k_cw8.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench.json 2> $O/bench.err
ls -la $O
