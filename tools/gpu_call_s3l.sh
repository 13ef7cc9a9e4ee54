#!/bin/bash
# session-3 call L: full GPU suite (RDF rewrite), RDF bench, LOVO end to end after the host fix
set -u
O=gpurun_out/s3l; mkdir -p $O
( timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 ) > $O/pytest_gpu.log
timeout 300 python tools/rdf_bench.py > $O/rdf_bench.json 2> $O/rdf_bench.err
timeout 300 python tools/rdf_bench.py 100000 5000 500 > $O/rdf_bench_big.json 2>> $O/rdf_bench.err
timeout 300 python tools/lovo_bench.py 5000 50000 > $O/lovo_1gpu.json 2> $O/lovo_1gpu.err
ls $O
