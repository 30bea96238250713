"""Known-answer test of the oracle's start-phase rule, setPhase<SynchronizePhase::Yes> (include/gen.crtp.h:127-156; rt.resynth
instantiates its synth with it, source/rt.resynth.lib.cpp:201-210): a NoteOn whose angle increment equals that of an
RT-active channel (a releasing one included) continues that channel's oscillator phase instead of drawing a random one.

The signal: a 6 kHz sine (8 samples per period, 32 periods per hop: every frame sees the same samples, so the analysis gives
bit-identical peaks frame after frame) with two tiny clicks. With max_track_pitches = 0 a note is only continued when its
pitch repeats exactly, so while a click sits in the window (16 frames) every frame releases the previous note and starts a new
one at a frequency that rounds to the same float; once the click has left the window the pre-click pitch is back. With a long
release all of these voices overlap. Known answer: because every new voice takes over the phase of the releasing ones, the
resynthesised signal is ONE sinusoid with a varying amplitude -- its phase against a 6 kHz quadrature reference never moves."""
import numpy as np


def sync_signal(amps=(0.003, 0.03), n=120000, sr=48000):
    """Stream 0 (click 0.003): every note-on lands on the float frequency of the sounding voices. Stream 1 (click 0.03): a
    few of them land one float away and start with a random phase -- a mix of both branches for the parity tests."""
    period = (0.5 * np.sin(2 * np.pi * np.arange(8) / 8)).astype(np.float32)  # tiled: the float32 samples repeat exactly
    xs = []
    for a in amps:
        x = np.tile(period, n // 8).astype(np.float64)
        x[20000] += a
        x[70000] += a
        xs.append(x)
    return np.stack(xs).astype(np.float32)


def quadrature_ratio(y, sr=48000, f=6000.0, block=256):
    """Largest |quadrature / in-phase| content over hop-sized blocks, the reference phase fitted on a steady stretch; and the
    smallest product of the signal with the in-phase reference (negative = a half period out of step)."""
    n = np.arange(len(y))
    w = 2 * np.pi * f / sr
    seg = slice(10000, 15000)
    phi = np.arctan2((y[seg] * np.cos(w * n[seg])).sum(), (y[seg] * np.sin(w * n[seg])).sum())
    ref, quad = np.sin(w * n + phi), np.cos(w * n + phi)
    m = (len(y) // block) * block
    ip = (y[:m] * ref[:m]).reshape(-1, block).sum(1)
    qp = (y[:m] * quad[:m]).reshape(-1, block).sum(1)
    on = ip > 0.05 * ip.max()
    return np.abs(qp[on] / ip[on]).max(), (y * ref).min() / np.abs(y).max()


def test_new_notes_continue_the_phase_of_sounding_channels(orc):
    x = sync_signal()
    env = orc.make_ahdsr(0, 0, 0, 20000, 1.0)
    orc.sync_count()
    out, stats = orc.roundtrip(x[:1], env=env, max_track_pitches=0.0)
    assert stats[0, 0] == 35 and stats[0, 1] == 0   # 1 + (16 click frames + 1) x 2 note-ons
    assert orc.sync_count() == 34                   # every note-on but the first found a channel to follow
    ratio, lowest = quadrature_ratio(out[0, :, 0])
    assert ratio < 0.01 and lowest > -1e-6, (ratio, lowest)   # one sinusoid: nothing in quadrature, never out of step
    orc.roundtrip(x[1:], env=env, max_track_pitches=0.0)
    assert 20 <= orc.sync_count() < 34              # stream 1 has note-ons on other float frequencies as well
    # the same run without the rule: random start phases, the voices beat against each other
    orc.set_sync_phase(False)
    try:
        out2, stats2 = orc.roundtrip(x[:1], env=env, max_track_pitches=0.0)
    finally:
        orc.set_sync_phase(True)
    assert np.array_equal(stats, stats2) and orc.sync_count() == 0
    ratio2, lowest2 = quadrature_ratio(out2[0, :, 0])
    assert ratio2 > 1.0 and lowest2 < -0.5
    assert np.abs(out - out2).max() > 0.5


def test_no_synchronisation_on_distinct_frequencies(orc):
    # a harmonic tone: 8 partials, all different, tracked from the first frame on -> the rule never fires
    sr, n = 48000, 24000
    t = np.arange(n)
    x = sum((1.0 / h) * np.sin(2 * np.pi * 220.0 * h * t / sr + h) for h in range(1, 9))
    x = (0.9 * x / np.abs(x).max()).astype(np.float32)[None, :]
    orc.sync_count()
    orc.roundtrip(x)
    assert orc.sync_count() == 0
