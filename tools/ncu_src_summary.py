#!/usr/bin/env python
"""Summarise `ncu --page source --csv` output: per-instruction samples, top stall sites, totals per opcode."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = collections.Counter()
for r in data:
    for s in stalls:
        agg[s] += int(r[ix[s]] or 0)
print("total samples", tot)
print("stall totals:", {k: v for k, v in agg.most_common(10)})
byop = collections.Counter(); execop = collections.Counter()
for r in data:
    op = r[ix["Source"]].split()[0] if r[ix["Source"]] else "?"
    if op.startswith("@"): op = r[ix["Source"]].split()[1]
    byop[op.split(".")[0]] += int(r[ix["# Samples"]] or 0)
    execop[op.split(".")[0]] += int(r[ix["Instructions Executed"]] or 0)
print("samples by opcode:", byop.most_common(12))
print("warp-inst executed by opcode:", execop.most_common(14), "total", sum(execop.values()))
top = sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[: int(sys.argv[2]) if len(sys.argv) > 2 else 25]
for r in top:
    st = {s[6:]: int(r[ix[s]] or 0) for s in stalls if int(r[ix[s]] or 0) > 0}
    print(r[ix["Address"]][-5:], r[ix["# Samples"]].rjust(5), r[ix["Instructions Executed"]].rjust(8), r[ix["Source"]][:60].ljust(60), dict(sorted(st.items(), key=lambda kv: -kv[1])[:3]))
