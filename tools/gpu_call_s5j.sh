#!/bin/bash
# 2 GPUs: the driver's scaling leg at N=2 (both arms) + sharded LOVO / all-pairs / RDF check
set -u
O=gpurun_out/s5j; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err
timeout 600 $TR --master-port 29532 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > $O/bench_ref_n2.json 2> $O/bench_ref_n2.err
timeout 600 $TR --master-port 29533 tools/multi_gpu_check.py > $O/multi_gpu_check.log 2>&1
timeout 300 $TR --master-port 29534 tools/lovo_bench.py 5000 50000 > $O/lovo_2gpu.log 2>&1
