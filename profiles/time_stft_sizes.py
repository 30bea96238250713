"""Analysis throughput for windows other than 4096 (the reference default 0.1814 s is 8192 points at 44.1 kHz and 16384
after zero padding at 48 kHz): python profiles/time_stft_sizes.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
pkg = g.load_package()
import cpp_audio_b200.workloads as wl
ctx = pkg.Context(0)
dev = torch.device("cuda:0")
S, sec = 32, 60.0
x = torch.from_numpy(wl.tiled_streams(S, sec)).to(dev)
for W, hop in [(1024, 256), (2048, 256), (4096, 256), (8192, 256), (8708, 4320), (16384, 512), (32768, 1024)]:
    an = pkg.SqMagAnalyzer(ctx, W, hop, max_peaks=32)
    out = an.allocate(S, x.shape[1], dev)
    for _ in range(2):
        an(x, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        an(x, out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    frames = S * an.num_frames(x.shape[1])
    nfft = pkg.fft_length_for(W, 1)
    alg = frames * (4 * hop + 4 * (nfft // 2 + 1))
    print(f"window {W:6d} hop {hop:5d} n_fft {nfft:6d}: {ms:8.3f} ms  {frames/ms/1e3:8.2f} M frames/s  {alg/ms/1e6:8.1f} GB/s algorithmic")
    del out, an
