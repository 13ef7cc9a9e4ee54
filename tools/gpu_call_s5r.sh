#!/bin/bash
# session-5 call H: what the driver runs at round end (tests, smoke, both bench arms) + launch list + full capture of the bench kernel
set -u
O=gpurun_out/s5r; mkdir -p $O
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) > $O/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
( time python bench.py --impl reference --steps 3 --warmup 1 ) > $O/bench_reference.json 2> $O/bench_reference.err
( time python bench.py ) > $O/bench.json 2> $O/bench.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_launches.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:super_cluster -s 3 -c 1 -o $O/bench_prof -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_full.log 2>&1
timeout 300 python tools/small_sweep.py > $O/small_sweep.jsonl 2> $O/small_sweep.err
timeout 300 python tools/misc_bench.py > $O/misc.json 2> $O/misc.err
timeout 300 python tools/lovo_bench.py > $O/lovo_e2e.json 2> $O/lovo_e2e.err
timeout 300 python tools/config1_bench.py > $O/config1.json 2> $O/config1.err
ls -la $O
