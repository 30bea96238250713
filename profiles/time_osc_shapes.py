"""Config 3 rendered with every CTA shape of the per-voice oscillator-bank launch (CA_OSC_SHAPE = warps per CTA, CA_OSC_CPW =
chunks per warp): which one is the default in oscbank.cu (not a bench number)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
pkg = g.load_package()
import cpp_audio_b200.workloads as wl
ctx = pkg.Context(0)
dev = torch.device("cuda:0")
V, ns = 1024, 480000
b = {k: torch.from_numpy(v).to(dev) for k, v in wl.c3_bank(V, 64).items()}
bank = pkg.SineBank(ctx, V, 64)
out = torch.empty((V, ns), dtype=torch.float32, device=dev)
env = pkg.Ahdsr(0, 0, 0, 0, 0, 0, 0, 1.0)
def run(n):
    for _ in range(n):
        bank.render(ns, b["f0_hz"], b["volume"], b["partial_volume"], b["partial_phase"], env, out=out)
ref = None
for path, shape, cpw in [(pa, sh, c) for pa, shapes in (("tensor", ("6", "8", "12", "16")), ("fp32", ("6", "12", "16"))) for sh in shapes for c in ("1", "2", "4", "8")]:
    if True:
        os.environ["CA_OSC_PATH"], os.environ["CA_OSC_SHAPE"], os.environ["CA_OSC_CPW"] = path, shape, cpw
        run(2)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(5); e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        chk = out[::97, ::1001].double().abs().sum().item()
        ref = chk if ref is None else ref
        print(f"{path:6s} warps/CTA {shape:>2}  chunks/warp {cpw}: {ms:.3f} ms  {V*ns/ms/1e6:.1f} G voice-samples/s  checksum {chk:.6f} (dev {abs(chk-ref):.2e})", flush=True)
