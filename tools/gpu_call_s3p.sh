#!/bin/bash
# session-3 call P: what the driver runs at round end (smoke, both bench arms) + launch list + full capture of the bench kernel
set -u
O=gpurun_out/s3p; mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
( time python bench.py --impl reference --steps 3 --warmup 1 ) > $O/bench_reference.json 2> $O/bench_reference.err
( time python bench.py ) > $O/bench.json 2> $O/bench.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_launches.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:super_cluster -s 3 -c 1 -o $O/bench_prof -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_full.log 2>&1
ls -la $O
