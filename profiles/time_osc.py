"""Timing of the oscillator-bank kernel on config 3 (not a bench number)."""
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
for _ in range(3):
    bank.render(ns, b["f0_hz"], b["volume"], b["partial_volume"], b["partial_phase"], env, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    bank.render(ns, b["f0_hz"], b["volume"], b["partial_volume"], b["partial_phase"], env, out=out)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"osc C3 {ms:.3f} ms  {V*ns/ms/1e6:.1f} G voice-samples/s  {V*ns*64/ms/1e9:.2f} T partial-samples/s {sys.argv[1] if len(sys.argv)>1 else ''}")
