#!/bin/bash
set -u
O=gpurun_out/s5c; mkdir -p $O
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > $O/pytest.log
{
echo "lovo 50kx5k";   timeout 120 python tools/one.py 5000 50000 lovo 0 0 0 5
echo "lovo 100kx10k"; timeout 120 python tools/one.py 10000 100000 lovo 0 0 0 3
echo "fit 100kx1500"; timeout 120 python tools/one.py 1500 100000 fit 0 0 0 5
echo "fit 100kx3000"; timeout 120 python tools/one.py 3000 100000 fit 0 0 0 5
echo "fit 50kx5k";    timeout 120 python tools/one.py 5000 50000 fit 0 0 0 5
echo "wb 100kx10k";   timeout 120 python tools/one.py 10000 100000 wb 0 0 0 3
echo "wb 100kx2000 (K=8 resident)";   timeout 120 python tools/one.py 2000 100000 wb 0 0 0 3
echo "lovo 100kx2000 (K=8 resident)"; timeout 120 python tools/one.py 2000 100000 lovo 0 0 0 3
} > $O/one.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err
timeout 600 python tools/large_frames_bench.py > $O/large_frames.json 2> $O/large_frames.err
