#!/bin/bash
# session-3 call C: moments stream kernel (tests + bench), LOVO pass ncu after the rework
set -u
O=gpurun_out/s3c; mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
for shape in "10000 100000" "1500 100000" "20000 50000" "3000 100000"; do timeout 300 python tools/moments_bench.py $shape >> $O/moments.jsonl 2>> $O/moments.err; done
timeout 300 python tools/one.py 5000 50000 lovo 2 0 1 2 > $O/one_lovo_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:super_cluster -s 1 -c 1 -o $O/lovo_prof -f python tools/one.py 5000 50000 lovo 2 0 1 2 > $O/ncu_lovo.log 2>&1
timeout 300 python tools/moments_bench.py 10000 100000 > $O/mom_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:moments_stream -s 2 -c 1 -o $O/moments_prof -f python tools/moments_bench.py 10000 100000 > $O/ncu_moments.log 2>&1
ls -la $O
