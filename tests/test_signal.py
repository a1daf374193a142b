"""Event-table front-end: oracle vs the unmodified mnormalized (CPU), GPU vs oracle (bit-exact float32)."""
import numpy as np
import pytest

from deepmod_b200 import synth
from oracle import ref_harness, signal_ref


def _oracle(raw_off, raw, ev_off, start, length):
    means, stdvs = [], []
    for r in range(len(raw_off) - 1):
        x = raw[raw_off[r]:raw_off[r + 1]]
        st, ln = start[ev_off[r]:ev_off[r + 1]], length[ev_off[r]:ev_off[r + 1]]
        sig = signal_ref.normalize(x, st, ln)
        m, s = signal_ref.event_stats(sig, st, ln)
        means.append(m); stdvs.append(s)
    return np.concatenate(means), np.concatenate(stdvs)


@pytest.mark.skipif(not ref_harness.available(), reason="reference tree not mounted")
def test_normalize_equals_unmodified_mnormalized():
    raw_off, raw, ev_off, start, length = synth.make_raw_signals(3, seed=3, mean_events=400)
    for r in range(3):
        x = raw[raw_off[r]:raw_off[r + 1]]
        st, ln = start[ev_off[r]:ev_off[r + 1]].astype(np.uint64), length[ev_off[r]:ev_off[r + 1]].astype(np.uint64)
        assert np.array_equal(signal_ref.normalize(x, st, ln), signal_ref.normalize_reference(x, st, ln))


def test_numpy_summation_model():
    """The GPU reproduces numpy's pairwise summation; this pins that model against numpy itself."""
    def pw(a):
        n = len(a)
        if n < 8:
            res = 0.0
            for v in a:
                res += v
            return res
        if n <= 128:
            r = [a[j] for j in range(8)]
            i = 8
            while i < n - (n % 8):
                for j in range(8):
                    r[j] += a[i + j]
                i += 8
            res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
            while i < n:
                res += a[i]
                i += 1
            return res
        n2 = n // 2
        n2 -= n2 % 8
        return pw(a[:n2]) + pw(a[n2:])
    rng = np.random.default_rng(1)
    for _ in range(500):
        a = np.round(rng.normal(0, 1.5, int(rng.integers(1, 700))), 3)
        assert pw(list(a)) / len(a) == np.mean(a)


@pytest.mark.gpu
def test_gpu_event_stats_bit_exact():
    from deepmod_b200 import capi, checkpoint
    raw_off, raw, ev_off, start, length = synth.make_raw_signals(24, seed=5)
    want_m, want_s = _oracle(raw_off, raw, ev_off, start, length)
    with capi.Context(checkpoint.random_model(0), 0) as ctx:
        m, s = ctx.event_stats(raw_off, raw, ev_off, start, length)
        ms = ctx.last_timing()[1]
    bad = int((m != want_m).sum() + (s != want_s).sum())
    print("event stats: %d events, %d samples, %d mismatching values, %.3f ms" % (len(m), len(raw), bad, ms))
    assert np.abs(m - want_m).max() <= 1.001e-3 and np.abs(s - want_s).max() <= 1.001e-3
    assert bad == 0


@pytest.mark.gpu
def test_gpu_event_stats_truncated_raw():
    """A raw array shorter than its event table says (truncated file): numpy slicing clips the normalisation span
    and the last events (myDetect.py:266-270, :334-343); the GPU must do the same instead of reading out of bounds."""
    from deepmod_b200 import capi, checkpoint
    raw_off, raw, ev_off, start, length = synth.make_raw_signals(6, seed=9, mean_events=300)
    # cut the tail of reads 1 and 4: their last events now reach past the end of the raw array
    keep, new_off = [], [0]
    for r in range(6):
        x = raw[raw_off[r]:raw_off[r + 1]]
        if r in (1, 4):
            last = ev_off[r + 1] - 1
            x = x[:int(start[last - 3] + length[last - 3] // 2)]
        keep.append(x)
        new_off.append(new_off[-1] + len(x))
    raw2, raw_off2 = np.concatenate(keep), np.array(new_off, np.int64)
    want_m, want_s = [], []
    for r in range(6):
        x = raw2[raw_off2[r]:raw_off2[r + 1]]
        st, ln = start[ev_off[r]:ev_off[r + 1]], length[ev_off[r]:ev_off[r + 1]]
        with np.errstate(all="ignore"):
            sig = signal_ref.normalize(x, st, ln)
            m, s = signal_ref.event_stats(sig, st, ln)
        want_m.append(m); want_s.append(s)
    want_m, want_s = np.concatenate(want_m), np.concatenate(want_s)
    with capi.Context(checkpoint.random_model(0), 0) as ctx:
        m, s = ctx.event_stats(raw_off2, raw2, ev_off, start, length)
    ok = ~np.isnan(want_m)              # events entirely past the end: mean of an empty slice (nan in numpy)
    assert ok.sum() > len(ok) - 12
    assert np.array_equal(m[ok], want_m[ok]) and np.array_equal(s[ok], want_s[ok])
