#!/usr/bin/env python
"""Top warp-stall sites of one kernel from `ncu -i X.ncu-rep --page source --csv --kernel-name ... > file.csv`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]; idx = {k: i for i, k in enumerate(hdr)}
data = [r for r in rows[h + 1:] if len(r) == len(hdr)]
def n(r, k):
    try: return int(r[idx[k]] or 0)
    except ValueError: return 0
tot = sum(n(r, "# Samples") for r in data)
print("total samples", tot, "instructions", len(data))
stalls = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
agg = {k: sum(n(r, k) for r in data) for k in stalls}
print("by reason:", sorted(((v, k) for k, v in agg.items() if v), reverse=True)[:8])
for r in sorted(data, key=lambda r: -n(r, "# Samples"))[: int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    st = sorted(((n(r, k), k[6:]) for k in stalls), reverse=True)[:2]
    print(str(n(r, "# Samples")).rjust(6), r[idx["Address"]][-5:], r[idx["Source"]][:100].ljust(100), st)
