#!/bin/bash
# compute-sanitizer over small instances of every kernel family (race detection / sanitizers).
#   bash tools/sanitize.sh [outdir]        on a machine with a B200
# memcheck on everything; racecheck (shared-memory hazards) and synccheck on the kernels that hand data between warps
# through shared memory.  Each tool run is bounded by its own timeout.
set -u
O=${1:-gpurun_out/sanitize}; mkdir -p $O
CS=/usr/local/cuda/bin/compute-sanitizer
run() {  # name, tool, timeout, families...
  local name=$1 tool=$2 to=$3; shift 3
  timeout $to $CS --tool $tool --error-exitcode 9 --print-limit 20 python tools/sanitize_target.py "$@" > $O/$name.log 2>&1
  echo "$name rc=$? $(grep -c '^ok' $O/$name.log) families-ok; $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $O/$name.log | tail -n 1)" | tee -a $O/summary.txt
}
: > $O/summary.txt
python tools/sanitize_target.py > $O/plain.log 2>&1; echo "plain rc=$? $(tail -n 1 $O/plain.log)" | tee -a $O/summary.txt
run memcheck_super memcheck 600 small mid cluster generic gather split misc
run memcheck_other memcheck 600 allpairs moments rdf
run racecheck_super racecheck 600 small mid cluster generic
run racecheck_other racecheck 600 allpairs moments rdf
run synccheck_super synccheck 400 small mid cluster generic gather split
