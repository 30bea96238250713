"""Static SASS instruction mix of the frame loop of stft4096p_kernel<12,1,1,1,1> in an object file: python profiles/loop_mix.py stft.o"""
import collections, re, subprocess, sys
obj = sys.argv[1]
names = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
fn = next(l.split()[-1] for l in names.splitlines() if "Function :" in l and "stft4096p_kernelILi12ELi1ELb1ELb1ELb1" in l)
sass = subprocess.run(["cuobjdump", "-sass", "-fun", fn, obj], capture_output=True, text=True).stdout
ins = []
for l in sass.splitlines():
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
back = []
for a, t in ins:
    m = re.search(r"BRA(?:\.\w+)*\s+(?:[!U]*P?\w*,\s*)?0x([0-9a-f]+)", t)
    if m and int(m.group(1), 16) < a:
        back.append((a, int(m.group(1), 16)))
hi, lo = max(back, key=lambda x: x[0] - x[1])
c = collections.Counter()
for a, t in ins:
    if lo <= a <= hi:
        op = re.sub(r"^@!?U?P\d+\s+", "", t).split()[0]
        c[op if op.startswith(("LDS", "STS", "LDG", "STG")) else op.split(".")[0]] += 1
print("loop", hex(lo), hex(hi), "instructions", sum(c.values()))
print(", ".join(f"{k} {v}" for k, v in c.most_common(30)))
