"""Top SASS instructions of an ncu report by stall samples, with shared-memory wavefronts and global tag requests:
python profiles/top_sass.py report.ncu-rep [n]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = raw.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.reader(io.StringIO("\n".join(lines[start:]))))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
def num(r, h):
    try: return float(r[ix[h]])
    except Exception: return 0.0
body = [r for r in rows[1:] if len(r) == len(hdr)]
tot = sum(num(r, "# Samples") for r in body)
print("total samples", tot, "instructions", len(body))
print("shared wavefronts", sum(num(r, "L1 Wavefronts Shared") for r in body), "ideal", sum(num(r, "L1 Wavefronts Shared Ideal") for r in body),
      "global tag requests", sum(num(r, "L1 Tag Requests Global") for r in body))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(num(r, h) for r in body) for h in stalls}
print("stall reasons:", ", ".join(f"{h[6:]} {100*v/tot:.1f}%" for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for idx, r in sorted(enumerate(body), key=lambda t: -num(t[1], "# Samples"))[:n]:
    top = max(stalls, key=lambda h: num(r, h))
    print(f"{idx:5d} {100*num(r,'# Samples')/tot:5.1f}%  shw {num(r,'L1 Wavefronts Shared'):9.0f} gtag {num(r,'L1 Tag Requests Global'):8.0f}  {top[6:]:12s} {r[ix['Source']].strip()[:90]}")
