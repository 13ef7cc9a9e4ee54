#!/bin/bash
# session-3 call J: all-pairs GEMM pipeline shapes (BK x stages), RDF ncu
set -u
O=gpurun_out/s3j; mkdir -p $O
( timeout 600 python -m pytest tests -m gpu -x -q -k "allpairs" 2>&1 | tail -3 ) > $O/pytest_allpairs.log
CUBLAS_REF=1 timeout 300 python tools/allpairs_bench.py 20000 3000 > $O/ap_32_2.json 2>&1
CUBLAS_REF=0 timeout 300 python tools/allpairs_bench.py 20000 3000 fit > $O/ap_32_2_fit.json 2>&1
for v in 16_3 32_3; do
  GCB_LIB=$PWD/gochem_b200/lib/libgochem_b200_ap$v.so CUBLAS_REF=0 timeout 300 python tools/allpairs_bench.py 20000 3000 > $O/ap_$v.json 2>&1
  GCB_LIB=$PWD/gochem_b200/lib/libgochem_b200_ap$v.so CUBLAS_REF=0 timeout 300 python tools/allpairs_bench.py 20000 3000 fit > $O/ap_${v}_fit.json 2>&1
done
timeout 300 python tools/rdf_bench.py 30000 1500 500 > $O/rdf_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rdf_kernel -s 1 -c 1 -o $O/rdf_prof -f python tools/rdf_bench.py 30000 1500 500 > $O/ncu_rdf.log 2>&1
ls $O
