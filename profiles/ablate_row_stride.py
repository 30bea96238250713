"""K2 (stft4096p_kernel) against the row stride of the |X|^2 output: 2049 floats (rows packed, every 128-byte store of a warp
straddles two lines) ... 2080 (rows start on 128-byte lines). Timing only (not a bench number); the library variant is chosen
with CA_B200_LIB (profiles/build_variant.sh)."""
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
rows = full["sqmag"].shape[0]
ref = None
for stride in (2049, 2052, 2056, 2064, 2080, 2112):
    buf = torch.zeros((rows, stride), dtype=torch.float32, device=dev)
    out = dict(full)
    out["sqmag"] = buf[:, :2049]
    for _ in range(3):
        an(x, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        an(x, out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    if ref is None:
        ref = (out["sqmag"].clone(), out["peak_bin"].clone(), out["peak_count"].clone())
        same = True
    else:
        same = bool(torch.equal(out["sqmag"], ref[0]) and torch.equal(out["peak_bin"], ref[1]) and torch.equal(out["peak_count"], ref[2]))
    print(f"lib {os.path.basename(os.environ.get('CA_B200_LIB', 'default'))} row stride {stride}: {ms:.3f} ms  {rows / ms / 1e3:.1f} Mframes/s  same as 2049: {same}", flush=True)
    del buf, out
