#!/bin/bash
# session-3 call I (8 GPUs): RDF tests + bench on one GPU, then multi-GPU check incl. sharded RDF, streaming with NUMA-local pinned rings
set -u
O=gpurun_out/s3i; mkdir -p $O
( timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 ) > $O/pytest_gpu.log
timeout 300 python tools/rdf_bench.py > $O/rdf_bench.json 2> $O/rdf_bench.err
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
AP_FRAMES=2000 AP_ATOMS=300 timeout 600 $TR --master-port 29511 tools/multi_gpu_check.py > $O/multi_gpu_check.log 2>&1
timeout 300 $TR --master-port 29514 tools/stream_bench.py 125000 20000 f32soa > $O/stream_f32soa_8gpu_numa.log 2>&1
GCB_NUMA_BIND=0 timeout 300 $TR --master-port 29515 tools/stream_bench.py 125000 20000 f32soa > $O/stream_f32soa_8gpu_nobind.log 2>&1
timeout 300 $TR --master-port 29516 tools/stream_bench.py 62500 20000 f64aos > $O/stream_f64aos_8gpu_numa.log 2>&1
nvidia-smi topo -m > $O/topo.txt 2>&1; lscpu | grep -i "numa\|socket\|model name" > $O/cpu.txt 2>&1
ls $O
