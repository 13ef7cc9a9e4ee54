#!/bin/bash
# One profiling pass on a GPU box (1 GPU): plain runs first, then ncu.  Summaries land in gpurun_out/; copy what is to be judged
# into profiles/.   bash tools/profile_round.sh r2
tag=${1:-r2}
out=gpurun_out
mkdir -p $out
summ() {  # rep-file kernel-regex -> key metrics as text
  ncu -i $1 --page raw --csv > $out/.raw.csv 2>/dev/null
  python - $out/.raw.csv <<'PY'
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
if len(rows) < 3:
    print("no data"); sys.exit()
hdr, units = rows[0], rows[1]
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]
for r in rows[2:]:
    for k in keys:
        if k in hdr:
            i = hdr.index(k)
            print("%-92s %-14s %s" % (k, units[i], r[i][:110]))
    print()
PY
}
# ---- 1. bench: plain run, then the launch list of the same command (skip with SKIP_BENCH=1) ----
if [ -z "$SKIP_BENCH" ]; then
python bench.py --steps 20 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
python bench.py --steps 5 --warmup 3 --no-extra --no-cpu-baseline > $out/${tag}_bench_short.json 2>/dev/null &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv python bench.py --steps 5 --warmup 3 --no-extra --no-cpu-baseline > $out/${tag}_ncu_launches.log 2>&1
python - $out/${tag}_launches.csv > $out/${tag}_launches_summary.txt <<'PY'
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]; k = hdr.index("Kernel Name"); v = hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    try: t = float(r[v].replace(",", ""))
    except ValueError: continue
    a = agg.setdefault(r[k][:110], [0, 0.0]); a[0] += 1; a[1] += t
tot = sum(a[1] for a in agg.values())
print("launch list of `python bench.py --steps 5 --warmup 3 --no-extra --no-cpu-baseline` (first 400 launches, ncu gpu__time_duration.sum, ns; cold-cache, serialised)")
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%6d launches %14.0f ns %6.2f %%  %s" % (n, t, 100 * t / tot, name))
PY
cat $out/${tag}_launches_summary.txt | head -12
fi
# ---- 2. the dominant kernels, one full capture each ----
python tools/tune.py 10000 100000 fit 0,0,0,-1 > $out/${tag}_plain_fit.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:super_cluster -s 3 -c 1 -f -o $out/${tag}_fit python tools/tune.py 10000 100000 fit 0,0,0,-1 > /dev/null 2>&1
summ $out/${tag}_fit.ncu-rep > $out/${tag}_fit_ncu_summary.txt
python tools/tune.py 10000 100000 wb 0,0,0,-1 > $out/${tag}_plain_wb.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:super_cluster -s 3 -c 1 -f -o $out/${tag}_wb python tools/tune.py 10000 100000 wb 0,0,0,-1 > /dev/null 2>&1
summ $out/${tag}_wb.ncu-rep > $out/${tag}_wb_ncu_summary.txt
python tools/tune.py 5000 50000 lovo 0,0,0,-1 > $out/${tag}_plain_lovo.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:super_cluster -s 3 -c 1 -f -o $out/${tag}_lovo python tools/tune.py 5000 50000 lovo 0,0,0,-1 > /dev/null 2>&1
summ $out/${tag}_lovo.ncu-rep > $out/${tag}_lovo_ncu_summary.txt
python tools/allpairs_bench.py 8192 3000 > $out/${tag}_plain_allpairs.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:allpairs_dmma_kernel -c 1 -f -o $out/${tag}_allpairs python tools/allpairs_bench.py 8192 3000 > /dev/null 2>&1
summ $out/${tag}_allpairs.ncu-rep > $out/${tag}_allpairs_ncu_summary.txt
python tools/allpairs_bench.py 8192 3000 fit > $out/${tag}_plain_allpairs_fit.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:allpairs_fit_dmma_kernel -c 1 -f -o $out/${tag}_allpairs_fit python tools/allpairs_bench.py 8192 3000 fit > /dev/null 2>&1
summ $out/${tag}_allpairs_fit.ncu-rep > $out/${tag}_allpairs_fit_ncu_summary.txt
rm -f $out/${tag}_wb.ncu-rep $out/${tag}_allpairs_fit.ncu-rep $out/${tag}_fit.ncu-rep   # keep the reports that are read at source level
for f in fit wb lovo allpairs allpairs_fit; do echo "== $f"; grep -E "gpu__time_duration|dram__bytes_read.sum |dram__bytes_write.sum |dmma_cycles|pipe_fp64|issue_active|registers" $out/${tag}_${f}_ncu_summary.txt; done
tail -q -n 1 $out/${tag}_plain_*.log | cut -c1-300
