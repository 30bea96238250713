"""SASS census of the built library: per kernel, how often the mnemonics that matter appear (cuobjdump -sass, no GPU needed).
usage: python profiles/sass_census.py [kernel-name substrings ...] > profiles/sass_census_r02.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "cpp.audio_b200", "libcpp_audio_b200.so")
want = sys.argv[1:] or ["stft4096p_kernel", "osc_umma_kernel", "oscbank_kernel", "rfft_warp_kernel", "rfft_cta_kernel", "rfft_cluster2_kernel", "big2_",
                        "stft_cta_kernel", "stft_warp_kernel",
                        "stft64_peaks_kernel", "pitch_track3_kernel", "pitch_prefix_kernel", "tracks_kernel", "sound_engine_kernel"]
MNEMONICS = ["UTMALDG", "UBLKCP", "FFMA2", "FADD2", "FMUL2", "FFMA", "FADD", "FMUL", "DFMA", "DADD", "DMUL", "MUFU", "LDS", "STS", "LDG", "STG",
             "SHFL", "BAR", "SYNCS", "UTCHMMA", "UTCQMMA", "HMMA", "LDTM", "UTCBAR", "UCGABAR_ARV", "UCGABAR_WAIT", "ST", "STAS", "STL", "LDL"]
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
cur, counts, total = None, {}, {}
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        counts[cur] = collections.Counter()
        total[cur] = 0
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur:
        op = m.group(1)
        total[cur] += 1
        for k in MNEMONICS:
            if op == k or (k in ("LDS", "STS", "LDG", "STG", "SHFL", "BAR", "MUFU", "SYNCS") and op.startswith(k)):
                counts[cur][k] += 1
print(f"cuobjdump -sass {os.path.relpath(so, ROOT)} (sm_100a): static instruction counts per kernel. Tensor-core mnemonics are expected in the "
      "oscillator bank only (UTCHMMA / LDTM / UTCBAR in osc_umma_kernel = tcgen05.mma / tcgen05.ld / tcgen05.commit, HMMA in oscbank_kernel = "
      "mma.sync); UCGABAR_* / ST (generic store to a mapa address) / STAS (st.async) mark the cluster kernels' exchange through "
      "distributed shared memory; UTMALDG / UBLKCP are TMA tensor and bulk copies")
for name in sorted(counts):
    if not any(w in name for w in want):
        continue
    c = counts[name]
    short = re.sub(r"\(anonymous namespace\)::|ca::", "", name)[:110]
    print(f"\n{short}\n  {total[name]} instructions: " + ", ".join(f"{k} {c[k]}" for k in MNEMONICS if c[k]))
