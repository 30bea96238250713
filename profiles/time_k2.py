"""Time stft4096p_kernel on config 2 x 32 streams (spectrum + peaks), 3 rounds of 20 launches; the library variant is chosen with
CA_B200_LIB (profiles/build_variant.sh). Prints a checksum of the outputs so that variants can be compared. Not a bench number."""
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
out = an.allocate(S, x.shape[1], dev)
rows = out["sqmag"].shape[0]
best = 1e9
for rnd in range(3):
    for _ in range(3):
        an(x, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        an(x, out)
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 20)
cs = (float(out["sqmag"].double().sum()), int(out["peak_bin"].long().sum()), int(out["peak_count"].long().sum()), float(out["peak_freq"].double().nan_to_num().sum()))
print(f"lib {os.path.basename(os.environ.get('CA_B200_LIB', 'default'))}: {best:.4f} ms  {rows / best / 1e3:.1f} Mframes/s  checksum {cs}", flush=True)
