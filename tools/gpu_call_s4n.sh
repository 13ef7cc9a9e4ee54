#!/bin/bash
set -u
O=gpurun_out/s4n; mkdir -p $O
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 ) > $O/pytest.log
timeout 600 python tools/small_sweep.py > $O/sweep_2_2.jsonl 2> $O/sweep.err
for v in 3_3 4_3; do GCB_LIB=$PWD/gochem_b200/lib/libgochem_b200_ss$v.so timeout 600 python tools/small_sweep.py > $O/sweep_$v.jsonl 2>> $O/sweep.err; done
