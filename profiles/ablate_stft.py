"""Timing ablation of the STFT kernel: which outputs cost what (not a bench number)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
pkg = g.load_package()
import cpp_audio_b200.workloads as wl
ctx = pkg.Context(0)
dev = torch.device("cuda:0")
S = 32
x = torch.from_numpy(wl.tiled_streams(S, 60.0)).to(dev)
an = pkg.SqMagAnalyzer(ctx, 4096, 256, max_peaks=32)
full = an.allocate(S, x.shape[1], dev)
combos = {"sqmag+peaks": full, "peaks only": {k: v for k, v in full.items() if k != "sqmag"}, "sqmag only": {"sqmag": full["sqmag"]}, "nothing": {}}
for name, out in combos.items():
    for _ in range(3):
        an(x, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        an(x, out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"variant {os.environ.get('CA_STFT_VARIANT','default')} {name:12s} {ms:.3f} ms  {S*11235/ms/1e3:.1f} Mframes/s")
