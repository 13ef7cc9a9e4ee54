#!/bin/bash
set -u
O=gpurun_out/s5z4; mkdir -p $O
( timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 ) > $O/pytest.log
timeout 200 python tools/large_frames_bench.py 100000 8000 > $O/large_u4.json 2> $O/err.log
GCB_LIB=$PWD/gochem_b200/lib/libgochem_b200_u8.so timeout 200 python tools/large_frames_bench.py 100000 8000 > $O/large_u8.json 2>> $O/err.log
timeout 100 python tools/split_small_probe.py > $O/small_u4.jsonl 2>> $O/err.log
GCB_LIB=$PWD/gochem_b200/lib/libgochem_b200_u8.so timeout 100 python tools/split_small_probe.py > $O/small_u8.jsonl 2>> $O/err.log
