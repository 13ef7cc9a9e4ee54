#!/bin/bash
set -u
O=gpurun_out/s5b; mkdir -p $O
timeout 300 python tools/subset_fit_probe.py 100000 20000 300 1000 3000 4096 > $O/probe_default.json 2> $O/probe.err
GCB_LIB=$PWD/gochem_b200/lib/libgochem_b200_mid1.so timeout 300 python tools/subset_fit_probe.py 100000 20000 300 1000 3000 4096 > $O/probe_mid1.json 2>> $O/probe.err
timeout 300 python tools/one_large.py 100000 4000 300 > $O/one_large.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"transform_frames_kernel|peratom_accum_kernel" -c 2 -o $O/large_prof python tools/one_large.py 100000 4000 300 > $O/ncu.log 2>&1
