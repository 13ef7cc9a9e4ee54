#!/bin/bash
set -u
O=gpurun_out/s5a; mkdir -p $O
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > $O/pytest.log
timeout 600 python tools/large_frames_bench.py > $O/large_frames.json 2> $O/large_frames.err
