"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count / total / average / share."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[start]
ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
agg = collections.OrderedDict()
tot = 0.0
n = 0
for r in rows[start + 1:]:
    if len(r) <= iv or r[im] != "gpu__time_duration.sum":
        continue
    name = r[ik].split("(")[0]
    us = float(r[iv].replace(",", "")) / 1e3
    unit = r[hdr.index("Metric Unit")]
    if unit in ("us", "usecond"):
        us *= 1e3
    elif unit in ("ms", "msecond"):
        us *= 1e6
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += us
    tot += us
    n += 1
print("launches %d total %.3f ms" % (n, tot / 1e3))
for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-70s n=%4d total %9.1f us avg %8.1f us share %5.1f%%" % (name[:70], c, t, t / c, 100 * t / tot))
