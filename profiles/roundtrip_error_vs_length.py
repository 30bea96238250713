"""Round-trip error against the double oracle as a function of the stream length, for the FP64 analysis (default) and the
float32 analysis (CA_RT_F32_ANALYSIS=1). Run once per mode; prints one line per length. Also times the 32 x 60 s call."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
pkg = g.load_package(); orc = g.load_oracle()
import cpp_audio_b200.workloads as wl
ctx = pkg.Context(0)
mode = "float32 analysis" if os.environ.get("CA_RT_F32_ANALYSIS") else "FP64 analysis"
S = 4
for sec in (1.0, 2.0, 5.0, 10.0, 20.0, 40.0, 60.0):
    x = wl.c4_streams(S, sec, first=17)
    out, stats = pkg.resynth_roundtrip(ctx, x)
    ref, rstats = orc.roundtrip(x, threads=orc.hardware_threads())
    err = np.abs(out - ref).max(axis=(1, 2))
    print(f"{mode}: {S} streams x {sec:5.1f} s: max-abs error per stream {' '.join(f'{e:.2e}' for e in err)}  stats equal: {np.array_equal(stats, rstats)}")
x0 = wl.tiled_streams(32, 60.0)
x = pkg.pinned_empty(x0.shape); x[:] = x0
obuf = pkg.pinned_empty(x0.shape + (2,))
pkg.resynth_roundtrip(ctx, x, out=obuf)
ts = []
for _ in range(5):
    t0 = time.perf_counter(); pkg.resynth_roundtrip(ctx, x, out=obuf); ts.append(time.perf_counter() - t0)
print(f"{mode}: 32 streams x 60 s round trip: {1e3 * min(ts):.1f} ms (best of 5)")
