"""Timing of the FFT vocoder (ca_vocoder_host) against the CPU oracle (not a bench number)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
pkg = g.load_package(); orc = g.load_oracle()
import cpp_audio_b200.workloads as wl
ctx = pkg.Context(0)
S, sec = int(sys.argv[1]) if len(sys.argv) > 1 else 8, float(sys.argv[2]) if len(sys.argv) > 2 else 60.0
mod0 = wl.tiled_streams(S, sec)
rng = np.random.default_rng(1)
mod = pkg.pinned_empty(mod0.shape); mod[:] = mod0
car = pkg.pinned_empty(mod0.shape); car[:] = rng.uniform(-0.5, 0.5, mod0.shape).astype(np.float32)
out = pkg.pinned_empty(mod0.shape)
pkg.vocoder(ctx, mod, car, out=out)                  # first call allocates the context's buffers
t0 = time.perf_counter(); pkg.vocoder(ctx, mod, car, out=out); t1 = time.perf_counter()
print(f"GPU vocoder (reference defaults): {S} streams x {sec} s: {t1-t0:.3f} s -> {S*mod.shape[1]/(t1-t0)/1e6:.1f} M samples/s")
nt = orc.hardware_threads(); n = min(S, nt)
t0 = time.perf_counter()
ref, _ = orc.vocoder(mod[:n], car[:n], modulator_window_seconds=0.10, stride_seconds=0.005, count_bands=5, max_freq=20000.0, threads=nt)
t1 = time.perf_counter()
print(f"CPU oracle: {n} streams x {sec} s on {nt} threads: {t1-t0:.3f} s -> {n*mod.shape[1]/(t1-t0)/1e6:.2f} M samples/s")
print("max err", np.abs(out[:n] - ref).max(), "peak", np.abs(ref).max())
