"""Aggregate an ncu gpu__time_duration launch list (csv) per kernel for the LAST proof in the log."""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
names = [r[ki] for r in rows[1:]]; vals = [float(r[vi].replace(',', '')) for r in rows[1:]]
last = max(i for i, n in enumerate(names) if n.startswith('tr_init'))
agg = collections.OrderedDict()
for n, v in zip(names[last:], vals[last:]):
    k = n.split('(')[0][:48]
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"total {tot/1e3:.1f} us over {sum(a[0] for a in agg.values())} launches (ncu: cold-cache, serialised; compare shares)")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:50s} {a[0]:5d} {a[1]/1e3:10.1f} us  {100*a[1]/tot:5.1f}%  avg {a[1]/a[0]/1e3:8.1f}")
