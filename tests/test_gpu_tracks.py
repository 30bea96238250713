"""GPU parity of ca_oscbank_render_tracks (a note schedule on the rt.resynth voice pool) against the oracle's VoicePool, which
steps the reference elements one sample at a time in double (include/gen.crtp.h:382-622, :350-378)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ON, OFF, CHANGE = 0, 1, 2


def schedule(seed, n_notes, n_samples, changes=6, step=256, f_lo=80.0, f_hi=9000.0):
    """Notes with random onsets; each gets a few NoteChanges on a `step` grid (glides and volume moves) and a NoteOff."""
    rng = np.random.default_rng(seed)
    ev = []
    for i in range(n_notes):
        t0 = int(rng.integers(0, n_samples // 2))
        f = float(np.exp(rng.uniform(np.log(f_lo), np.log(f_hi))))
        ev.append((t0, 100 + i, ON, f, float(rng.uniform(0.2, 1.0))))
        t = t0
        for _ in range(int(rng.integers(0, changes + 1))):
            t += step * int(rng.integers(1, 4))
            f *= float(2.0 ** rng.uniform(-0.08, 0.08))
            ev.append((t, 100 + i, CHANGE, f, float(rng.uniform(0.2, 1.0))))
        if rng.uniform() < 0.8:
            ev.append((t + int(rng.integers(1, 3000)), 100 + i, OFF, 0.0, 0.0))
    ev.sort(key=lambda e: e[0])   # stable: the events of a note stay in order
    return ev


def test_schedule_matches_the_reference_pool(pkg, orc, ctx):
    n = 48000
    ev = schedule(1, 40, n)
    out, stats = pkg.render_tracks(ctx, ev, n)
    ref, rstats = orc.render_tracks(ev, n)
    assert stats == rstats and stats[0] == 40 and stats[1] == 0
    assert np.abs(ref).max() > 0.05
    assert np.abs(out - ref).max() < 1e-5


def test_glide_shapes_and_envelopes(pkg, orc, ctx):
    # one note: up an octave over a long glide, back down over a short one, released in the middle of the second
    n = 30000
    ev = [(100, 1, ON, 440.0, 1.0), (4000, 1, CHANGE, 880.0, 0.5), (12000, 1, CHANGE, 330.0, 0.9), (12700, 1, OFF, 0, 0)]
    env = pkg.Ahdsr(600, 0, 200, 900, 0, 5000, 0, 0.6)
    for glide in (1024, 4320, 100):
        out, stats = pkg.render_tracks(ctx, ev, n, glide_samples=glide, env=env, stereo_spread=0.4)
        ref, rstats = orc.render_tracks(ev, n, glide_samples=glide, env=orc.make_ahdsr(600, 200, 900, 5000, 0.6), stereo_spread=0.4)
        assert stats == rstats == (1, 0)
        assert np.abs(out - ref).max() < 1e-5, glide
    assert np.abs(out[:100]).max() == 0.0


def test_same_frequency_note_on_continues_the_sounding_phase(pkg, orc, ctx):
    """setPhase<SynchronizePhase::Yes> (include/gen.crtp.h:127-156): note 2 starts at the increment note 1 is still releasing
    with and takes over its phase; note 3 has another frequency and starts with its own random phase; note 4 lands on the channel
    note 1 left, whose oscillator still holds note 1's increment (include/audioelement.h:2387-2388), and follows note 2."""
    n = 60000
    env = pkg.Ahdsr(0, 0, 0, 0, 0, 8000, 0, 1.0)
    ev = [(0, 1, ON, 1000.0, 0.8), (5000, 1, OFF, 0, 0), (5600, 2, ON, 1000.0, 0.8), (9000, 3, ON, 1500.0, 0.5),
          (30000, 4, ON, 2000.0, 0.7), (31000, 2, OFF, 0, 0), (33000, 5, ON, 1000.0, 0.3)]
    orc.sync_count()
    ref, rstats = orc.render_tracks(ev, n, env=orc.make_ahdsr(0, 0, 0, 8000, 1.0), mono=True)
    assert orc.sync_count() == 3      # notes 2, 4 (through the stale increment of its channel) and 5
    out, stats = pkg.render_tracks(ctx, ev, n, env=env, mono=True)
    assert stats == rstats == (5, 0)
    assert np.abs(out - ref).max() < 1e-5
    orc.set_sync_phase(False)
    try:
        nosync, _ = orc.render_tracks(ev, n, env=orc.make_ahdsr(0, 0, 0, 8000, 1.0), mono=True)
    finally:
        orc.set_sync_phase(True)
    assert np.abs(nosync - ref).max() > 0.05   # the case depends on the rule


def test_pool_limit_drops_and_reuses_channels(pkg, orc, ctx):
    n = 40000
    ev = schedule(7, 60, n, changes=2)
    out, stats = pkg.render_tracks(ctx, ev, n, max_channels=12)
    ref, rstats = orc.render_tracks(ev, n, max_channels=12)
    assert stats == rstats and stats[1] > 0
    assert np.abs(out - ref).max() < 1e-5


def test_device_output_and_mono(pkg, orc, ctx):
    n = 20000
    ev = schedule(3, 10, n)
    d = torch.empty((n,), dtype=torch.float32, device="cuda")
    out, stats = pkg.render_tracks(ctx, ev, n, mono=True, device_out=d)
    torch.cuda.synchronize()
    ref, rstats = orc.render_tracks(ev, n, mono=True)
    assert stats == rstats
    assert np.abs(out.cpu().numpy() - ref).max() < 1e-5
    host, _ = pkg.render_tracks(ctx, ev, n, mono=True)
    assert np.array_equal(host, out.cpu().numpy())


def test_edge_cases(pkg, ctx):
    out, stats = pkg.render_tracks(ctx, [], 5000)                       # no events: silence
    assert out.shape == (5000, 2) and np.abs(out).max() == 0 and stats == (0, 0)
    out, stats = pkg.render_tracks(ctx, [(10, 1, ON, 440.0, 1.0)], 0)   # nothing to render
    assert out.shape == (0, 2)
    out, stats = pkg.render_tracks(ctx, [(7000, 1, ON, 440.0, 1.0)], 5000)   # event beyond the end
    assert np.abs(out).max() == 0 and stats == (0, 0)
    with pytest.raises(Exception):
        pkg.render_tracks(ctx, [(10, 1, ON, 440.0, 1.0), (5, 1, OFF, 0, 0)], 100)   # unordered
    with pytest.raises(Exception):
        pkg.render_tracks(ctx, [(10, 1, 9, 440.0, 1.0)], 100)                          # unknown event type
    with pytest.raises(Exception):
        pkg.render_tracks(ctx, [], 100, env=pkg.Ahdsr(0, 5, 0, 0, 0, 0, 0, 1.0))      # eased envelope


def test_streaming_pool_matches_one_shot(pkg, ctx):
    """ca_tracks_open / post / pull: events posted in two batches, audio pulled in uneven blocks = the one-shot render."""
    import ctypes as C
    from cpp_audio_b200.synth import track_events
    n = 30000
    ev = schedule(11, 25, n)
    one, stats = pkg.render_tracks(ctx, ev, n)
    L = pkg.lib()
    d = pkg.TracksDesc(0, None, 0, 48000, 256, pkg.Ahdsr(0, 0, 0, 0, 0, 0, 0, 1.0), 254, 1.0, 1)
    h = C.c_void_p()
    ctx.check(L.ca_tracks_open(ctx.handle, C.byref(d), C.byref(h)), "ca_tracks_open")
    half = len(ev) // 2
    a1, a2 = track_events(ev[:half]), track_events(ev[half:])
    ctx.check(L.ca_tracks_post(h, C.cast(a1, C.c_void_p), half), "post")
    out = np.zeros((n, 2), np.float32)
    pos, posted = 0, False
    for blk in (1, 15, 1000, 4096, 33, 7000, 100000):
        if not posted and pos + blk > ev[half][0]:   # the second batch arrives before the pool reaches its first event
            ctx.check(L.ca_tracks_post(h, C.cast(a2, C.c_void_p), len(ev) - half), "post")
            posted = True
        take = min(blk, n - pos)
        if take <= 0:
            break
        buf = np.empty((take, 2), np.float32)
        ctx.check(L.ca_tracks_pull_host(h, take, buf.ctypes.data_as(C.c_void_p)), "pull")
        out[pos:pos + take] = buf
        pos += take
    st = (C.c_int64 * 5)()
    L.ca_tracks_state(h, st)
    L.ca_tracks_close(h)
    assert pos == n and st[0] == n and (st[1], st[2]) == stats
    # tile boundaries differ between the two runs, so the float32 volume-filter state is rounded at other places
    assert np.abs(out - one).max() < 2e-6
