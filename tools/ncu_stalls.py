"""summarise the source page of an .ncu-rep: stall samples by opcode / reason and the top stalled instructions
    python tools/ncu_stalls.py gpurun_out/x.ncu-rep [top_n]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr, data = rows[h], rows[h + 1:]
ix = {k: i for i, k in enumerate(hdr)}
stalls = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
byop, tot = collections.Counter(), collections.Counter()
for r in data:
    if len(r) < len(hdr):
        continue
    n = int(r[ix["# Samples"]] or 0)
    toks = [t for t in r[ix["Source"]].split() if not t.startswith("@")]
    byop[toks[0].split(".")[0] if toks else ""] += n
    for s in stalls:
        tot[s] += int(r[ix[s]] or 0)
print("samples", sum(byop.values()))
print("by opcode", byop.most_common(14))
print("by reason", tot.most_common(10))
for r in sorted((r for r in data if len(r) >= len(hdr)), key=lambda r: -int(r[ix["# Samples"]] or 0))[:top]:
    why = max(stalls, key=lambda s: int(r[ix[s]] or 0))
    print(f'{r[ix["# Samples"]]:>5} {r[0][-5:]} {r[ix["Source"]].strip()[:72]:72} x{r[ix["Instructions Executed"]]} {why}')
