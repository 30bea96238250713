"""Synthetic inputs of the shapes BASELINE.json names (SURVEY.md §8 d). numpy only; seeded; used by tests and bench.py."""
from __future__ import annotations

import numpy as np

SR = 48000


def c1_frames(batch: int = 4096, n: int = 1024, seed: int = 1234) -> np.ndarray:
    """C1: 1024-pt frames of an 8-sine mix, float32, peak <= 1."""
    rng = np.random.default_rng(seed)
    t = np.arange(n)
    k = 16 + 32 * np.arange(8)
    delta = rng.uniform(0, 1, (batch, 8))
    phi = rng.uniform(0, 2 * np.pi, (batch, 8))
    x = np.zeros((batch, n))
    for j in range(8):
        x += (0.5 / (j + 1)) * np.sin(2 * np.pi * (k[j] + delta[:, j:j + 1]) * t[None, :] / n + phi[:, j:j + 1])
    x /= np.maximum(1.0, np.abs(x).max(axis=1, keepdims=True))
    return x.astype(np.float32)


def harmonic_tone(n_samples: int, f0: float = 220.0, n_harm: int = 16, seed: int = 2345, sr: int = SR,
                  peak: float = 0.9) -> np.ndarray:
    """C2: sum_h (1/h) sin(2 pi f0 h n / sr + phi_h), normalised to `peak`."""
    rng = np.random.default_rng(seed)
    phi = rng.uniform(0, 2 * np.pi, n_harm)
    n = np.arange(n_samples, dtype=np.float64)
    x = np.zeros(n_samples)
    for h in range(1, n_harm + 1):
        x += (1.0 / h) * np.sin(2 * np.pi * f0 * h * n / sr + phi[h - 1])
    return (peak * x / np.abs(x).max()).astype(np.float32)


def c2_signal(seconds: float = 60.0) -> np.ndarray:
    return harmonic_tone(int(round(seconds * SR)))


def c4_streams(n_streams: int, seconds: float, first: int = 0, total: int = 256) -> np.ndarray:
    """C4: stream s has f0 = 110*2^(s/total), seed 4000+s. Returns float32 [n_streams, samples] for s in [first, first+n)."""
    n = int(round(seconds * SR))
    return np.stack([harmonic_tone(n, 110.0 * 2.0 ** ((first + s) / total), seed=4000 + first + s) for s in range(n_streams)])


def tiled_streams(n_streams: int, seconds: float) -> np.ndarray:
    """n_streams C4-style streams built from at most 8 distinct signals (cheap to generate for benchmarks)."""
    base = c4_streams(min(8, n_streams), seconds, total=8)
    reps = (n_streams + base.shape[0] - 1) // base.shape[0]
    return np.ascontiguousarray(np.tile(base, (reps, 1))[:n_streams])


def c3_bank(n_voices: int = 1024, n_partials: int = 64, seed: int = 3456):
    """C3: f0 = 55*2^(2v/V) Hz, partial volumes 1/i scaled to sum 1, phases U(-1,1), velocity 1 (volume = 0.1)."""
    rng = np.random.default_rng(seed)
    f0 = (55.0 * 2.0 ** (2.0 * np.arange(n_voices) / n_voices)).astype(np.float32)
    pv = 1.0 / np.arange(1, n_partials + 1)
    pv = np.tile((pv / pv.sum()).astype(np.float32), (n_voices, 1))
    pp = rng.uniform(-1, 1, (n_voices, n_partials)).astype(np.float32)
    volume = np.full(n_voices, 0.1, dtype=np.float32)
    return dict(f0_hz=f0, volume=volume, partial_volume=pv, partial_phase=pp)


def c5_frames(p: int, total_log2: int = 28) -> np.ndarray:
    """C5: batch = 2^total_log2 / N frames of U(-1,1), N = 2^p."""
    n = 1 << p
    batch = max(1, (1 << total_log2) // n)
    rng = np.random.default_rng(5000 + p)
    return rng.uniform(-1, 1, (batch, n)).astype(np.float32)
