"""Aggregate the per-instruction warp-stall samples of an ncu report (--set full --import-source on) over windows of W
SASS instructions:  python profiles/stall_regions.py report.ncu-rep [W]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
W = int(sys.argv[2]) if len(sys.argv) > 2 else 100
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
base = int(data[0][0], 16)
recs = [(int(r[0], 16) - base, r[1].strip(), int(r[ix["# Samples"]]), {c: int(r[ix[c]]) for c in cols},
         int(r[ix["Instructions Executed"]])) for r in data]
tot = sum(x[2] for x in recs)
print("total samples", tot, " kernel:", rows[0][1][:90])
allst = {c: sum(x[3][c] for x in recs) for c in cols}
print("all:", [(k[6:], round(100 * v / tot, 1)) for k, v in sorted(allst.items(), key=lambda t: -t[1])[:9]])
for s in range(0, len(recs), W):
    ch = recs[s:s + W]
    n = sum(x[2] for x in ch)
    if n == 0:
        continue
    st = {c: sum(x[3][c] for x in ch) for c in cols}
    ops = {}
    for x in ch:
        t = x[1].split()
        o = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        ops[o] = ops.get(o, 0) + 1
    top = sorted(ops.items(), key=lambda t: -t[1])[:4]
    tops = sorted(st.items(), key=lambda t: -t[1])[:4]
    print(f"{ch[0][0]:#7x} {100 * n / tot:5.1f}%  exec {max(x[4] for x in ch):>8}  {top}  {[(k[6:], v) for k, v in tops]}")
