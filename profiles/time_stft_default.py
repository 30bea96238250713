"""Time the fused analysis at the reference's default window (8708 / hop 4320 -> 16384 points, stft_cta_kernel<8>) on 32 streams x 60 s
and at 16384 / hop 512; CA_B200_LIB chooses a library variant. Prints a checksum of the outputs. Not a bench number."""
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
for W, hop in ((8708, 4320), (16384, 512)):
    an = pkg.SqMagAnalyzer(ctx, W, hop, max_peaks=32)
    out = an.allocate(S, x.shape[1], dev)
    best = 1e9
    for rnd in range(3):
        for _ in range(3):
            an(x, out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            an(x, out)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 10)
    rows = out["sqmag"].shape[0]
    cs = (float(out["sqmag"].double().sum()), int(out["peak_bin"].long().sum()), int(out["peak_count"].long().sum()))
    print(f"lib {os.path.basename(os.environ.get('CA_B200_LIB', 'default'))} window {W} hop {hop}: {best:.4f} ms  {rows / best / 1e3:.2f} Mframes/s  checksum {cs}", flush=True)
    del out, an
