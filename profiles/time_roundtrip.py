"""Timing of the analysis -> pitch pipeline -> resynthesis round trip (not a bench number)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
pkg = g.load_package(); orc = g.load_oracle()
import cpp_audio_b200.workloads as wl
ctx = pkg.Context(0)
S, sec = int(sys.argv[1]) if len(sys.argv) > 1 else 16, float(sys.argv[2]) if len(sys.argv) > 2 else 20.0
x0 = wl.tiled_streams(S, sec)
x = pkg.pinned_empty(x0.shape); x[:] = x0          # host audio in page-locked memory
obuf = pkg.pinned_empty(x0.shape + (2,))
pkg.resynth_roundtrip(ctx, x, out=obuf)             # first call allocates the context's staging buffers
t0 = time.perf_counter(); out, stats = pkg.resynth_roundtrip(ctx, x, out=obuf); t1 = time.perf_counter()
print(f"GPU path: {S} streams x {sec} s: {t1-t0:.3f} s -> {S*x.shape[1]/(t1-t0)/1e6:.1f} M output samples/s, notes {stats[:,0].sum()}")
t0 = time.perf_counter(); ref, _ = orc.roundtrip(x[:orc.hardware_threads()], threads=orc.hardware_threads()); t1 = time.perf_counter()
n = min(S, orc.hardware_threads())
print(f"CPU oracle: {n} streams x {sec} s on {orc.hardware_threads()} threads: {t1-t0:.3f} s -> {n*x.shape[1]/(t1-t0)/1e6:.2f} M output samples/s")
print("max err", np.abs(out[:n] - ref).max())
