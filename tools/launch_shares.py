"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list: launch_shares.py FILE [SKIP_REGEX]
(SKIP_REGEX: kernels left out of the total, e.g. the one-off SRS setup 'srs_')."""
import collections, csv, re, sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
skip = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    k = r[ki].split("(")[0]
    if skip and skip.search(k):
        continue
    v = float(r[vi].replace(",", ""))
    v = v / 1000 if r[ui] == "ns" else (v * 1000 if r[ui] == "ms" else v)
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"total {tot:.1f} us over {sum(a[0] for a in agg.values())} launches (ncu: cold-cache, serialised; compare shares)")
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{k[:54]:54s} {a[0]:5d} {a[1]:10.1f} us {100 * a[1] / tot:5.1f}%  avg {a[1] / a[0]:9.1f}")
