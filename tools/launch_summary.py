"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel name."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
H = rows[hdr]; ki = H.index('Kernel Name'); vi = H.index('Metric Value'); ui = H.index('Metric Unit')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 1:]:
  if len(r) <= vi: continue
  v = float(r[vi].replace(',', ''))
  v = v / 1e3 if r[ui] == 'ns' else v * 1e3 if r[ui] == 'ms' else v
  agg[r[ki][:110]][0] += 1; agg[r[ki][:110]][1] += v
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
  print("%10.1f us %5.1f%% n=%6d avg %8.2f us  %s" % (v[1], 100 * v[1] / tot, v[0], v[1] / v[0], k))
print("total %.2f ms over %d launches" % (tot / 1e3, sum(v[0] for v in agg.values())))
