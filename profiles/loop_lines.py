"""Instructions of the frame loop of stft4096p_kernel<12,1,1,1,1> per source line: python profiles/loop_lines.py stft.o [min]"""
import collections, os, re, subprocess, sys, tempfile
obj = os.path.abspath(sys.argv[1]); mn = int(sys.argv[2]) if len(sys.argv) > 2 else 6
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=d, capture_output=True)
cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
L = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cub)], capture_output=True, text=True).stdout.split("\n")
start = next(i for i, l in enumerate(L) if l.startswith("//----") and "stft4096p_kernelILi12ELi1ELb1ELb1ELb1" in l)
end = next(i for i in range(start + 5, len(L)) if L[i].startswith("//--------------------- .text."))
rows = []; cur = None
for l in L[start:end]:
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
    if m and cur:
        rows.append((int(m.group(1), 16), re.sub(r"^@!?U?P\d+\s+", "", m.group(2).strip()), cur))
back = []
for a, t, _ in rows:
    m = re.search(r"BRA(?:\.\w+)*\s+(?:[!U]*P?\w*,\s*)?`\(([^)]+)\)", t)
# loop bounds from cuobjdump (addresses agree)
sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
fn = next(l.split()[-1] for l in sass.splitlines() if "Function :" in l and "stft4096p_kernelILi12ELi1ELb1ELb1ELb1" in l)
s2 = subprocess.run(["cuobjdump", "-sass", "-fun", fn, obj], capture_output=True, text=True).stdout
bb = []
for l in s2.splitlines():
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
    if m:
        m2 = re.search(r"BRA(?:\.\w+)*\s+(?:[!U]*P?\w*,\s*)?0x([0-9a-f]+)", m.group(2))
        if m2 and int(m2.group(1), 16) < int(m.group(1), 16):
            bb.append((int(m.group(1), 16), int(m2.group(1), 16)))
hi, lo = max(bb, key=lambda x: x[0] - x[1])
cnt = collections.Counter(); per = collections.defaultdict(collections.Counter)
for a, t, c in rows:
    if lo <= a <= hi:
        cnt[c] += 1; per[c][t.split()[0].split(".")[0]] += 1
print("loop", hex(lo), hex(hi), sum(cnt.values()))
for c, v in sorted(cnt.items()):
    if v >= mn and c[0] != "packed.cuh":
        print(c[0], c[1], v, dict(per[c].most_common(5)))
print("packed.cuh", sum(v for c, v in cnt.items() if c[0] == "packed.cuh"), dict(sum((per[c] for c in cnt if c[0] == "packed.cuh"), collections.Counter()).most_common(8)))
