"""CPU known-answer tests of the optional pitch stages of RtResynth::step (source/rt.resynth.lib.cpp:1687-1707) as the oracle
restates them; expected values are derived by hand from the in-tree functions (algo.cpp:186-229, 307-371;
rt.resynth.lib.autotune.cpp:96-146, 189-218)."""
import numpy as np

MAJOR = [0., 2., 4., 5., 7., 9., 11., 12.]  # rt.resynth.lib.autotune.cpp:142-151


def test_defaults_are_the_identity(orc):
    p, v = orc.pitch_stages(orc.make_pitch_options(), [60.0, 64.0, 67.5], [0.1, 0.2, 0.3])
    assert (p == [60.0, 64.0, 67.5]).all() and (v == [0.1, 0.2, 0.3]).all()
    # ... except that entries closer than pitch_epsilon merge (autotune_pitches with the identity transform, algo.cpp:221-222)
    p, v = orc.pitch_stages(orc.make_pitch_options(), [60.0, 60.00005, 64.0], [0.1, 0.2, 0.3])
    assert (p == [60.0, 64.0]).all() and np.allclose(v, [0.3, 0.3], atol=1e-15)


def test_shift_pre_and_post(orc):
    p, v = orc.pitch_stages(orc.make_pitch_options(shift_pre=12.0, shift_post=-0.5), [60.0, 64.0], [0.1, 0.2])
    assert (p == [71.5, 75.5]).all() and (v == [0.1, 0.2]).all()


def test_harmonize_adds_to_an_existing_pitch_or_inserts(orc):
    # +7 on [60, 67, 72]: 67 exists -> its volume grows by v0; 74 is new and carries the ALREADY GROWN volume of 67 (the loop
    # reads pitches[i] after it was modified); 79 is new. The new entries are merged in by pitch (algo.cpp:333-359).
    p, v = orc.pitch_stages(orc.make_pitch_options(harmonize_pre=7.0), [60.0, 67.0, 72.0], [1.0, 2.0, 4.0])
    assert (p == [60.0, 67.0, 72.0, 74.0, 79.0]).all()
    assert (v == [1.0, 3.0, 4.0, 3.0, 4.0]).all()
    # negative amount: the new entries sort before the old ones
    p, v = orc.pitch_stages(orc.make_pitch_options(harmonize_post=-12.0), [60.0, 72.0], [1.0, 2.0])
    assert (p == [48.0, 60.0, 72.0]).all() and (v == [1.0, 3.0, 2.0]).all()


def test_autotune_closest_allowed_pitch(orc):
    # allowed [60, 62, 64], tolerance 0.6, max pitch 100 (algo.cpp:205-226; find_closest_pitch, autotune.cpp:189-218)
    o = orc.make_pitch_options(allowed_pitches=[60.0, 62.0, 64.0], max_pitch=100.0, tolerance=0.6)
    p, v = orc.pitch_stages(o, [59.5, 60.3, 61.0, 63.2, 64.1, 120.0], [1.0, 2.0, 4.0, 8.0, 16.0, 32.0])
    # 59.5 -> 60; 60.3 -> 60 (merged); 61.0 is equidistant: the lower neighbour wins (60), distance 1 > tolerance: dropped;
    # 63.2 -> 64, distance 0.8: dropped; 64.1 -> 64; 120 is above max_pitch: played as it is
    assert (p == [60.0, 64.0, 120.0]).all()
    assert (v == [3.0, 16.0, 32.0]).all()
    # an empty table silences everything below max_pitch (the transform returns no pitch, rt.resynth.lib.cpp:1821-1826)
    o = orc.make_pitch_options(allowed_pitches=np.zeros(0), max_pitch=100.0, tolerance=50.0)
    p, v = orc.pitch_stages(o, [50.0, 120.0], [1.0, 2.0])
    assert (p == [120.0]).all() and (v == [2.0]).all()


def test_autotune_musical_scale(orc):
    # C major around C4 = 60: 61.2 -> 62 (1.2 lies in [1, 3)); 58.9 is folded up an octave (70.9 -> B = 71) -> 59;
    # 60.99 -> 60 (below the midpoint 1); 66.0 -> 3-rd midpoint is 4.5, 6 = midpoint(5,7) is not < 6 -> 67
    o = orc.make_pitch_options(scale=MAJOR, scale_root_pitch=60.0, max_pitch=151.0, tolerance=2.0)
    p, v = orc.pitch_stages(o, [58.9, 60.99, 61.2, 66.0], [1.0, 2.0, 4.0, 8.0])
    assert np.allclose(p, [59.0, 60.0, 62.0, 67.0], atol=1e-12)
    assert (v == [1.0, 2.0, 4.0, 8.0]).all()


def test_all_stages_in_order(orc):
    # shift +1, harmonize +12, autotune to even pitches, shift -2, harmonize +4 (rt.resynth.lib.cpp:1687-1707)
    o = orc.make_pitch_options(shift_pre=1.0, harmonize_pre=12.0, shift_post=-2.0, harmonize_post=4.0,
                               allowed_pitches=np.arange(0.0, 128.0, 2.0), max_pitch=127.0, tolerance=1.0)
    p, v = orc.pitch_stages(o, [59.2, 63.1], [1.0, 2.0])
    # shift: 60.2, 64.1; harmonize: + 72.2, 76.1; autotune: 60, 64, 72, 76; shift: 58, 62, 70, 74;
    # harmonize +4: 62 exists (+1 -> 3); 66 is new (vol 3, the grown one) and is appended, so the vector is now
    # [58, 62, 70, 74 | 66]; looking for 70 + 4 = 74 the bisection of lower_bound probes 70, then the appended 66, and runs
    # off the end: the "closest" entry is 66, not the 74 that exists, and a SECOND 74 is appended (the precondition of
    # find_closest_pitch does not hold inside the loop, algo.cpp:335-348); 78 likewise. Merge: first run first on ties.
    assert (p == [58.0, 62.0, 66.0, 70.0, 74.0, 74.0, 78.0]).all()
    assert (v == [1.0, 3.0, 3.0, 1.0, 2.0, 1.0, 2.0]).all()


def test_round_trip_with_shift_moves_the_notes(orc):
    sr = 48000
    n = np.arange(sr // 2)
    x = (0.3 * np.sin(2 * np.pi * 440.0 * n / sr)).astype(np.float32)
    a, sa = orc.roundtrip(x)
    b, sb = orc.roundtrip(x, opts=orc.make_pitch_options(shift_pre=12.0))
    assert sa[0, 0] == 1 and sb[0, 0] == 1

    def dominant(y):
        spec = np.abs(np.fft.rfft(y[0, 8192:8192 + 8192, 0] * np.hanning(8192)))
        return np.argmax(spec) * sr / 8192

    assert abs(dominant(a) - 440.0) < 6 and abs(dominant(b) - 880.0) < 6
    c, sc = orc.roundtrip(x, opts=orc.make_pitch_options(harmonize_post=7.0))
    assert sc[0, 0] == 2  # the fifth is played as a second note
