timeout 300 python tools/step_breakdown.py rnnt > gpurun_out/rnnt_breakdown.txt 2>&1; cat gpurun_out/rnnt_breakdown.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_wsj.csv python tools/step_breakdown.py wsj > gpurun_out/ncu_wsj.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/launches_wsj.csv')) if len(r) > 10]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[vi].replace(',', ''))
    if r[ui] == 'ns': v /= 1e3
    elif r[ui] == 'ms': v *= 1e3
    elif r[ui] == 's' or r[ui]=='second': v *= 1e6
    a = agg.setdefault(r[ki][:50], [0, 0.0]); a[0] += 1; a[1] += v
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
    print("%-52s n=%5d total %9.1f us  avg %7.2f us" % (k, n, us, us / n))
PY
