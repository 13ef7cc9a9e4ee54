#!/bin/bash
# session-3 call K (2 GPUs): LOVO end to end (1 and 2 ranks)
set -u
O=gpurun_out/s3k; mkdir -p $O
timeout 300 python tools/lovo_bench.py 5000 50000 > $O/lovo_1gpu.json 2> $O/lovo_1gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/lovo_bench.py 5000 50000 > $O/lovo_2gpu.log 2>&1
ls $O
