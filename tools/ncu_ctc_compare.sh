cd $GRAFT_REPO_ROOT
NCU="ncu --metrics gpu__time_duration.sum --clock-control none --csv -c 40"
SL_CTC_FUSED=0 $NCU --log-file gpurun_out/ncu_ctc_two.csv tools/selftest ctc_bench > /dev/null 2>&1
SL_CTC_DBG=1 $NCU --log-file gpurun_out/ncu_ctc_dbg1.csv tools/selftest ctc_bench > /dev/null 2>&1
SL_CTC_DBG=1 SL_CTC_GRAD_CTAS=148 $NCU --log-file gpurun_out/ncu_ctc_dbg1_148.csv tools/selftest ctc_bench > /dev/null 2>&1
$NCU --log-file gpurun_out/ncu_ctc_fused.csv tools/selftest ctc_bench > /dev/null 2>&1
for f in two dbg1 dbg1_148 fused; do echo "== $f"; python - gpurun_out/ncu_ctc_$f.csv <<'PY'
import csv, sys, collections
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
agg = collections.defaultdict(list)
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    if r[ui] in ("ns", "nsecond"): v /= 1e3
    elif r[ui] in ("ms", "msecond"): v *= 1e3
    agg[r[ki][:60]].append(v)
for k, v in agg.items(): print("  %-60s n=%d  median %.1f us  min %.1f" % (k, len(v), sorted(v)[len(v)//2], min(v)))
PY
done
