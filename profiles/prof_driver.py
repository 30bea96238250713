"""Small driver for ncu captures: runs one kernel family a few times on cuda:0 (never a bench number).
    python profiles/prof_driver.py stft [streams]     python profiles/prof_driver.py osc [voices]     ... rfft [log2n]
    ... roundtrip [streams]     ... vocoder [streams]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()
import cpp_audio_b200.workloads as wl  # noqa: E402

what = sys.argv[1]
ctx = pkg.Context(0)
dev = torch.device("cuda:0")
if what == "stft":
    S = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    x = torch.from_numpy(wl.tiled_streams(S, 60.0)).to(dev)
    an = pkg.SqMagAnalyzer(ctx, 4096, 256, max_peaks=32)
    out = an.allocate(S, x.shape[1], dev)
    for _ in range(4):
        an(x, out)
elif what == "stft16k":  # the reference's default window: 0.1814 s / 0.09 s at 48 kHz -> 8708 samples, 16384-point frames, hop 4320
    S = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    x = torch.from_numpy(wl.tiled_streams(S, 60.0)).to(dev)
    an = pkg.SqMagAnalyzer(ctx, 8708, 4320, max_peaks=32)
    out = an.allocate(S, x.shape[1], dev)
    for _ in range(4):
        an(x, out)
elif what == "osc":
    V = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    b = {k: torch.from_numpy(v).to(dev) for k, v in wl.c3_bank(V, 64).items()}
    bank = pkg.SineBank(ctx, V, 64)
    out = torch.empty((V, 480000), dtype=torch.float32, device=dev)
    for _ in range(4):
        bank.render(480000, b["f0_hz"], b["volume"], b["partial_volume"], b["partial_phase"], pkg.Ahdsr(0, 0, 0, 0, 0, 0, 0, 1.0), out=out)
elif what == "rfft":
    p = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    x = torch.from_numpy(wl.c5_frames(p, 26)).to(dev)
    for _ in range(4):
        pkg.rfft_batched(ctx, x)
elif what == "roundtrip":
    S = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    x = wl.tiled_streams(S, 60.0)
    for _ in range(3):
        pkg.resynth_roundtrip(ctx, x)
elif what == "roundtrip_small":   # the FP64 analysis kernel (stft64) with its neighbours: 8 streams x 10 s
    x = wl.tiled_streams(8, 10.0)
    for _ in range(3):
        pkg.resynth_roundtrip(ctx, x)
elif what == "vocoder":
    S = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    rng = np.random.default_rng(1)
    mod = wl.tiled_streams(S, 60.0)
    car = rng.uniform(-0.5, 0.5, mod.shape).astype(np.float32)
    for _ in range(3):
        pkg.vocoder(ctx, mod, car)
torch.cuda.synchronize()
print("done", ctx.kernel_launches)
