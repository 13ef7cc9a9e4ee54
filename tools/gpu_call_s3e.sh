#!/bin/bash
# session-3 call E (8 GPUs): all-pairs matrix with peer-store tiles, LOVO/allreduce check, weak-scaling bench, streaming (config 5)
set -u
O=gpurun_out/s3e; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tools/multi_gpu_check.py > $O/multi_gpu_check.log 2>&1
timeout 600 $TR --master-port 29512 tools/allpairs_sharded_bench.py 20000 3000 > $O/allpairs_sharded_8gpu.log 2>&1
timeout 600 $TR --master-port 29513 tools/allpairs_sharded_bench.py 20000 3000 fit > $O/allpairs_sharded_fit_8gpu.log 2>&1
timeout 600 $TR --master-port 29514 tools/stream_bench.py 125000 20000 f32soa > $O/stream_f32soa_8gpu.log 2>&1
timeout 600 $TR --master-port 29515 tools/stream_bench.py 62500 20000 f64aos > $O/stream_f64aos_8gpu.log 2>&1
timeout 600 $TR --master-port 29516 bench.py --gpus 8 --steps 10 --warmup 3 > $O/bench_8gpu.json 2> $O/bench_8gpu.err
ls -la $O
