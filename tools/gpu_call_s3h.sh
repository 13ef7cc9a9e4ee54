#!/bin/bash
# session-3 call H: two-CTAs-per-SM variant of the LOVO pass
set -u
O=gpurun_out/s3h; mkdir -p $O
timeout 300 python tools/tune.py 5000 50000 lovo 2,0,1,1 > $O/tune_mb1.log 2>&1
GCB_TUNE_MB=2 timeout 300 python tools/tune.py 5000 50000 lovo 4,0,1,1 4,0,2,1 5,0,1,1 5,0,2,1 6,0,2,1 7,0,2,1 8,0,2,1 > $O/tune_mb2.log 2>&1
GCB_TUNE_MB=2 timeout 300 python tools/tune.py 10000 100000 lovo 8,0,1,1 8,0,2,1 > $O/tune_mb2_10k.log 2>&1
GCB_TUNE_MB=2 timeout 600 python -m pytest tests -m gpu -x -q -k "lovo or dev_sum or peratom" > $O/pytest_mb2.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench.json 2> $O/bench.err
ls $O
