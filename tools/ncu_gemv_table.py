"""per-launch table of an `ncu --set full` report of the GEMV kernels: time, DRAM bytes / throughput, issue, LSU, warps
    python tools/ncu_gemv_table.py report.ncu-rep [label ...]"""
import csv
import io
import subprocess
import sys

rep, labels = sys.argv[1], sys.argv[2:]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}


def col(r, name):
    return float(r[ix[name]].replace(",", "")) if name in ix and r[ix[name]] not in ("", "n/a") else float("nan")


want = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "MB read"), ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"), ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU %"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU %"), ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"), ("launch__registers_per_thread", "regs"),
        ("sm__cycles_active.avg", "active cyc"), ("sm__cycles_elapsed.max", "elapsed cyc")]
print("| launch | kernel | grid x block | " + " | ".join(w[1] for w in want) + " |")
print("|---|---|---|" + "---|" * len(want))
for i, r in enumerate(data):
    name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "")
    vals = []
    for m, _ in want:
        v = col(r, m)
        if m == "dram__bytes_read.sum" and units[ix[m]].lower().startswith("byte"):
            v /= 1e6
        if m == "dram__bytes_read.sum" and units[ix[m]].lower().startswith("kbyte"):
            v /= 1e3
        vals.append(f"{v:.2f}" if v == v else "-")
    lab = labels[i] if i < len(labels) else str(i)
    print(f"| {lab} | {name} | {r[ix['Grid Size']].strip()} x {r[ix['Block Size']].strip()} | " + " | ".join(vals) + " |")
