#!/bin/bash
# 8 GPUs: the driver's scaling leg at N=8 (both arms) + sharded LOVO / all-pairs / RDF check on the final session-5 build
set -u
O=gpurun_out/s5v; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29541 bench.py --gpus 8 --steps 10 --warmup 3 > $O/bench_n8.json 2> $O/bench_n8.err
timeout 300 $TR --master-port 29542 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > $O/bench_ref_n8.json 2> $O/bench_ref_n8.err
timeout 300 $TR --master-port 29543 tools/multi_gpu_check.py > $O/multi_gpu_check.log 2>&1
timeout 200 $TR --master-port 29544 tools/lovo_bench.py 5000 50000 > $O/lovo_8gpu.log 2>&1
