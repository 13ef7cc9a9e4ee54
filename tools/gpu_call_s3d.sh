#!/bin/bash
# session-3 call D (2 GPUs): GPU tests, all-pairs single-GPU timing after the staged epilogue, 2-GPU peer-store all-pairs check + bench
set -u
O=gpurun_out/s3d; mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
CUBLAS_REF=0 timeout 300 python tools/allpairs_bench.py 20000 3000 > $O/allpairs_1gpu.json 2> $O/allpairs_1gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py > $O/multi_gpu_check.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/allpairs_sharded_bench.py 20000 3000 > $O/allpairs_sharded_2gpu.log 2>&1
ls -la $O
