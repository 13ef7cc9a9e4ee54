#!/bin/bash
# final check of round 1: what the driver runs at round end
set -u
O=gpurun_out/s5w; mkdir -p $O
( timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) > $O/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
( time python bench.py --impl reference --steps 3 --warmup 1 ) > $O/bench_reference.json 2> $O/bench_reference.err
( time python bench.py ) > $O/bench.json 2> $O/bench.err
timeout 120 python tools/config1_bench.py > $O/config1.json 2> $O/config1.err
