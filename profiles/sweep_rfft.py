"""BASELINE config 1 and config 5: batched real FFT throughput vs the HBM roofline (prints a table, one JSON line per size).
    python profiles/sweep_rfft.py [total_log2=28]"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
pkg = g.load_package()
ctx = pkg.Context(0)
dev = torch.device("cuda:0")
peak = 6451.5
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
total_log2 = int(sys.argv[1]) if len(sys.argv) > 1 else 28
rows = []
for p in list(range(10, 21)):
    n = 1 << p
    batch = max(1, (1 << total_log2) // n)
    x = torch.empty((batch, n), dtype=torch.float32, device=dev).uniform_(-1, 1)
    for _ in range(3):
        out = pkg.rfft_batched(ctx, x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        out = pkg.rfft_batched(ctx, x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    bytes_alg = batch * (4 * n + 8 * (n // 2 + 1))
    gbs = bytes_alg / (ms * 1e-3) / 1e9
    # independent cross-check of the numbers (torch.fft = cuFFT, never the implementation)
    ref = torch.fft.rfft(x[:2].double(), dim=1)
    err = (out[:2] - ref).abs().max().item() / (n ** 0.5)
    rows.append(dict(n_fft=n, batch=batch, ms=ms, frames_per_s=batch / (ms * 1e-3), gbs=gbs, frac_hbm=gbs / peak, err=err))
    print(json.dumps(rows[-1]), flush=True)
    del x, out
# config 1: 1024-pt, batch 4096 (33.6 MB: L2 resident, launch-latency dominated) through a CUDA-graph-free loop of 100 launches
x = torch.empty((4096, 1024), dtype=torch.float32, device=dev).uniform_(-1, 1)
for _ in range(10):
    pkg.rfft_batched(ctx, x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(100):
    pkg.rfft_batched(ctx, x)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 100
print(json.dumps(dict(config="C1", n_fft=1024, batch=4096, ms=ms, frames_per_s=4096 / (ms * 1e-3), gbs=4096 * 8200 / (ms * 1e-3) / 1e9,
                      note="33.6 MB working set: L2 resident, per-launch latency included")))
