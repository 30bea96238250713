"""Stress check of the cluster FFT kernels (compute-sanitizer is not available on this pool): full GPUs' worth of frames, several
repetitions while other streams keep the SMs busy, EVERY output frame compared with torch.fft in double (an independent anchor, not
the implementation). A missed barrier or a store landing after its row was read shows up as a sporadic mismatch.
    python profiles/stress_fft_cluster.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
pkg = g.load_package(); ctx = pkg.Context(0)
side = torch.cuda.Stream()
noise = torch.empty((1 << 26,), device="cuda")
worst = {}
for p, batch in ((15, 4099), (16, 2051), (17, 1027)):
    n = 1 << p
    x = torch.empty((batch, n), device="cuda").uniform_(-1, 1)
    ref = torch.fft.rfft(x.double(), dim=1)
    scale = 1.0 / np.sqrt(n)
    w = 0.0
    for rep in range(6):
        with torch.cuda.stream(side):  # competing work on another stream: clusters get scheduled next to foreign CTAs
            for _ in range(4):
                noise.uniform_(-1, 1)
        X = pkg.rfft_batched(ctx, x)
        torch.cuda.synchronize()
        err = ((X - ref).abs().amax(dim=1) * scale)
        w = max(w, err.max().item())
        bad = (err > 1e-5).nonzero()
        assert bad.numel() == 0, (p, rep, bad[:8].flatten().tolist())
    worst[p] = w
    del x, ref, X
print("cluster FFT stress ok, worst normalised error per length:", worst)
