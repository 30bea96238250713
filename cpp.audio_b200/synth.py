"""Host side of the synthesis path (sine oscillator bank), torch tensors in/out, compute in libcpp_audio_b200.so.

`SineBank` plays the role of a pool of VolumeAdjusted<MultiEnveloped<SineOscillatorAlgo,AHDSREnvelope>> elements driven by
ImplCRTP (include/gen.crtp.h:350-378, :382-531): notes are described up front (frequency, velocity, start phase, key delay,
release time, harmonics) and rendered for the whole duration instead of 16-frame blocks.
"""
from __future__ import annotations

import ctypes as C

import torch

from ._capi import Ahdsr, Context, OscBankDesc, lib

MIX_NONE, MIX_STEREO, MIX_MONO = 0, 1, 2
BASE_VOLUME_SINE = 0.1  # reduceUnadjustedVolumes * soundBaseVolume(SINE), include/audioelement.h:2370,1270; include/sound.h:84-85


def _opt_ptr(t):
    return None if t is None else t.data_ptr()


class SineBank:
    def __init__(self, ctx: Context, max_voices: int, max_partials: int, sample_rate: int = 48000):
        self.ctx = ctx
        self.sample_rate = sample_rate
        self.max_voices = max_voices
        self.max_partials = max_partials
        ctx.check(lib().ca_oscbank_plan(ctx.handle, max_voices, max_partials), "ca_oscbank_plan")

    def render(self, n_samples: int, f0_hz: torch.Tensor, volume: torch.Tensor, partial_volume: torch.Tensor,
               partial_phase: torch.Tensor, env, start_angle=None, key_delay=None, release_at=None, pan=None,
               mix: int = MIX_NONE, out: torch.Tensor | None = None, first_sample: int = 0) -> torch.Tensor:
        """All tensors on the same CUDA device. env: one Ahdsr or a list of Ahdsr (one per voice)."""
        dev = f0_hz.device
        nv, npart = partial_volume.shape
        for t, dt in ((f0_hz, torch.float32), (volume, torch.float32), (partial_volume, torch.float32),
                      (partial_phase, torch.float32)):
            if not (t.is_cuda and t.dtype == dt and t.is_contiguous()):
                raise ValueError("note parameters must be contiguous CUDA tensors of the documented dtype")
        envs = [env] if isinstance(env, Ahdsr) else list(env)
        arr = (Ahdsr * len(envs))(*envs)
        env_host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
        env_dev = env_host.to(dev)
        if mix == MIX_NONE:
            shape = (nv, n_samples)
        elif mix == MIX_STEREO:
            shape = (n_samples, 2)
        else:
            shape = (n_samples,)
        if out is None:
            out = torch.empty(shape, dtype=torch.float32, device=dev)
        d = OscBankDesc(nv, npart, n_samples, self.sample_rate, f0_hz.data_ptr(), volume.data_ptr(), _opt_ptr(start_angle),
                        _opt_ptr(key_delay), _opt_ptr(release_at), partial_volume.data_ptr(), partial_phase.data_ptr(),
                        env_dev.data_ptr(), len(envs), _opt_ptr(pan), mix, first_sample)
        self.ctx.check(lib().ca_oscbank_render(self.ctx.handle, C.byref(d), out.data_ptr(),
                                               torch.cuda.current_stream(dev).cuda_stream), "ca_oscbank_render")
        self._keep = (env_dev,)  # the render is asynchronous: keep the parameter block alive
        return out


def render_bank(ctx: Context, n_samples: int, f0_hz, volume, partial_volume, partial_phase, env, sample_rate=48000, **kw):
    bank = SineBank(ctx, partial_volume.shape[0], partial_volume.shape[1], sample_rate)
    return bank.render(n_samples, f0_hz, volume, partial_volume, partial_phase, env, **kw)
