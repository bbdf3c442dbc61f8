#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, mean us, share."""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.OrderedDict()
tot = 0.0
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")[:80]
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v; tot += v
print("%d launches, %.1f us total" % (sum(a[0] for a in agg.values()), tot))
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print("%9.1f us avg  x%4d  %5.1f%%  %s" % (t / c, c, 100 * t / tot, k))
