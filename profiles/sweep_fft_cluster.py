"""Cluster (single-pass, DSMEM transposition) vs the previous kernels for N = 2^14 ... 2^17: error against torch.fft in double and
time over 1 GiB of input. One subprocess per setting (the switches are read once per process).
    python profiles/sweep_fft_cluster.py            # all settings
    python profiles/sweep_fft_cluster.py worker     # (internal)"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "worker":
    import torch
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    pkg = g.load_package()
    ctx = pkg.Context(0)
    peak = 6451.5
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    for p in [int(a) for a in sys.argv[2:]]:
        n = 1 << p
        batch = (1 << 28) // n
        x = torch.empty((batch, n), dtype=torch.float32, device="cuda").uniform_(-1, 1)
        for _ in range(3):
            out = pkg.rfft_batched(ctx, x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            out = pkg.rfft_batched(ctx, x)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        sel = [0, 1, batch // 2, batch - 1]
        ref = torch.fft.rfft(x[sel].double(), dim=1)
        err = (out[sel] - ref).abs().max().item() / (n ** 0.5)
        gbs = batch * (4 * n + 8 * (n // 2 + 1)) / (ms * 1e-3) / 1e9
        print(json.dumps(dict(n_fft=n, cluster=os.environ.get("CA_FFT_CLUSTER"), cg=os.environ.get("CA_FFT_CG"), ms=round(ms, 4),
                              gbs=round(gbs, 1), frac_hbm=round(gbs / peak, 3), err=err)), flush=True)
        del x, out
    sys.exit(0)
for cluster, cg, sizes in [("0", "0", "14 15 16 17"), ("2", "2", "14 15 16 17"), ("2", "4", "15 16 17"), ("3", "0", "14 15 16 17")]:
    env = dict(os.environ, CA_FFT_CLUSTER=cluster, CA_FFT_CG=cg)
    subprocess.run([sys.executable, os.path.abspath(__file__), "worker"] + sizes.split(), env=env, check=False)
