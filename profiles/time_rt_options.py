"""Round trip (32 streams x 60 s, pinned host buffers) with and without the optional pitch stages (not a bench number)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
pkg = g.load_package()
import cpp_audio_b200.workloads as wl
ctx = pkg.Context(0)
S, sec = 32, 60.0
x0 = wl.tiled_streams(S, sec)
x = pkg.pinned_empty(x0.shape); x[:] = x0
obuf = pkg.pinned_empty(x0.shape + (2,))
MAJOR = [0., 2., 4., 5., 7., 9., 11., 12.]
cases = {
    "no options": None,
    "shift -12, harmonize +12 (notes doubled)": pkg.pitch_options(shift_pre=-12.0, harmonize_post=12.0),
    "autotune: C major scale": pkg.pitch_options(scale=MAJOR, scale_root_pitch=60.0, max_pitch=151.0, tolerance=1.0),
    "autotune: chord table (33 pitches) + harmonize +7": pkg.pitch_options(
        allowed_pitches=np.sort(np.array([72.0 + i + 12 * o for o in range(-5, 6) for i in (0, 4, 7)])), max_pitch=120.0,
        tolerance=2.0, harmonize_post=7.0),
}
for name, opts in cases.items():
    pkg.resynth_roundtrip(ctx, x, out=obuf, opts=opts)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); out, stats = pkg.resynth_roundtrip(ctx, x, out=obuf, opts=opts); ts.append(time.perf_counter() - t0)
    print(f"{name:52s} {min(ts)*1e3:6.1f} ms   notes started {stats[:,0].sum():6d}, dropped {stats[:,1].sum()}")
