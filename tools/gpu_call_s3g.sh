#!/bin/bash
# session-3 call G (1 GPU): RDF tests + timing, full GPU suite, bench line
set -u
O=gpurun_out/s3g; mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 600 python tools/rdf_bench.py > $O/rdf_bench.json 2> $O/rdf_bench.err
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err
timeout 300 python tools/stream_bench.py 60000 20000 f32soa > $O/stream1.log 2>&1
ls -la $O
