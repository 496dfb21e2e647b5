"""Per-source-line warp-stall sample shares of one kernel from an .ncu-rep captured with --import-source on.
usage: python tools/ncu_lines.py report.ncu-rep [top_n]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
files = {}
cur = None
hdr = None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1]
        files[cur] = []
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if cur is not None and hdr is not None and r and r[0] != "":
        try:
            files[cur].append((int(r[0]), int(r[hdr.index("# Samples")]), int(r[hdr.index("Instructions Executed")]), r[1]))
        except ValueError:
            pass
tot = sum(s for v in files.values() for _, s, _, _ in v)
print("total samples", tot)
allv = [(f.split("/")[-1], ln, s, i, src) for f, v in files.items() for ln, s, i, src in v]
byfile = {}
for f, ln, s, i, src in allv:
    byfile[f] = byfile.get(f, 0) + s
print({k: round(100 * v / tot, 1) for k, v in byfile.items()})
for f, ln, s, i, src in sorted(allv, key=lambda x: -x[2])[:top]:
    print(f"{f}:{ln:4d} {100 * s / tot:5.1f}% inst {i:10d}  {src.strip()[:120]}")
