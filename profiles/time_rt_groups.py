"""Round trip (32 streams x 60 s, pinned host buffers) for the stream-group count given by CA_RT_GROUPS (not a bench number)."""
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
pkg.resynth_roundtrip(ctx, x, out=obuf)
ts = []
for _ in range(5):
    t0 = time.perf_counter(); out, stats = pkg.resynth_roundtrip(ctx, x, out=obuf); ts.append(time.perf_counter() - t0)
print(f"CA_RT_GROUPS={os.environ.get('CA_RT_GROUPS', '4')}: best {min(ts)*1e3:.1f} ms, median {sorted(ts)[2]*1e3:.1f} ms, notes {stats[:,0].sum()}, checksum {float(np.abs(out).sum()):.6e}")
