"""Generates the golden fixtures in this directory FROM THE ORACLE (the reference publishes none and cannot be built
here: its arithmetic lives in the absent sibling repo cpp.algorithms). They freeze the oracle, so that a later change to
oracle/ that alters results is caught; they do not pin the oracle to upstream (oracle/SPEC.md: analysis parity unpinned).

Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

orc = g.load_oracle()
g.load_package()
import cpp_audio_b200.workloads as wl  # noqa: E402

here = os.path.dirname(os.path.abspath(__file__))

x = wl.c2_signal(0.25)  # 12000 samples -> 31 frames
r = orc.stft_analyze(x, 4096, 256)
np.savez_compressed(os.path.join(here, "analysis_c2_slice.npz"), x=x, peak_bin=r["peak_bin"][0, :, :16],
                    peak_freq=r["peak_freq"][0, :, :16], peak_db=r["peak_db"][0, :, :16],
                    sqmag_decimated=r["sqmag"][0][:, ::16])

b = wl.c3_bank(4, 16)
rel = np.array([3000, -1, 50, 4000], np.int64)
kd = np.array([0, 10, 100, 3], np.int32)
sa = np.array([0.0, 0.5, -0.3, 0.9], np.float32)
n = 6000
out = orc.oscbank_render(b["f0_hz"], b["volume"], b["partial_volume"], b["partial_phase"], n,
                         env=orc.make_ahdsr(200, 30, 400, 1000, sustain=0.7), release_at=rel, key_delay=kd, start_angle=sa)
np.savez_compressed(os.path.join(here, "synth_bank.npz"), n_samples=n, env_adhr=np.array([200, 30, 400, 1000]), sustain=0.7,
                    release_at=rel, key_delay=kd, start_angle=sa, out_decimated=out[:, ::7], **b)
print("golden fixtures written")
