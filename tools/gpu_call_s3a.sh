#!/bin/bash
# session-3 call A: re-validate the tree on a GPU, LOVO-pass tuning + ncu, all-pairs ncu
set -u
O=gpurun_out/s3a; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/box.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err
# LOVO pass shape sweep: default lib (8 consumer warps) and 12/16-warp builds
timeout 300 python tools/tune.py 5000 50000 lovo 1,0,1,1 2,0,1,1 2,0,2,1 3,0,1,1 3,0,2,1 4,0,2,1 > $O/tune_lovo_cw8.log 2>&1
for cw in 12 16; do
  GCB_LIB=$PWD/gochem_b200/lib/libgochem_b200_cw$cw.so timeout 300 python tools/tune.py 5000 50000 lovo 1,0,1,1 2,0,1,1 2,0,2,1 3,0,1,1 3,0,2,1 4,0,2,1 > $O/tune_lovo_cw$cw.log 2>&1
done
timeout 300 python tools/moments_bench.py 10000 100000 > $O/moments.log 2>&1
# ncu: LOVO pass (resident, C2), full set with source
timeout 300 python tools/one.py 5000 50000 lovo 2 0 1 2 > $O/one_lovo_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:super_cluster -s 1 -c 1 -o $O/lovo_prof -f python tools/one.py 5000 50000 lovo 2 0 1 2 > $O/ncu_lovo.log 2>&1
# ncu: all-pairs DMMA GEMM on a smaller matrix (8192 frames)
CUBLAS_REF=0 timeout 300 python tools/allpairs_bench.py 8192 3000 > $O/allpairs_plain.log 2>&1 &&
CUBLAS_REF=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:allpairs_dmma -s 1 -c 1 -o $O/allpairs_prof -f python tools/allpairs_bench.py 8192 3000 > $O/ncu_allpairs.log 2>&1
ls -la $O
