"""Host algebra of the filtfilt scan (csrc/filt_consts.cpp, CPU only): the cascade realisation the GPU runs has
the transfer function and the initial condition of scipy's single DF2T recurrence over (b, a, zi), is exact to
~1e-10 where scipy's own evaluation wobbles by ~3e-7, and its chunk carry is a plain fp64 mat-vec."""
import ctypes

import numpy as np
from scipy import signal

from oracle import ref_pipeline as rp
from speechsplit_b200 import _lib


def _cascade(chunk=256):
    lib = _lib.load()
    b, a = rp.butter_highpass(30, 16000, 5)
    zi = np.ascontiguousarray(signal.lfilter_zi(b, a))
    sec, zic, m = np.zeros(15), np.zeros(5), np.zeros(25)
    rc = lib.ssfe_filt_cascade(b.ctypes.data, a.ctypes.data, zi.ctypes.data, chunk, sec.ctypes.data, zic.ctypes.data,
                               m.ctypes.data)
    assert rc == 0
    return b, a, zi, sec.reshape(3, 5), zic, m.reshape(5, 5)


def _sos(sec, zic):
    sos = np.array([[s[0], s[1], s[2], 1.0, s[3], s[4]] for s in sec])
    zs = np.zeros((3, 2))
    zs[0, 0], zs[1], zs[2] = zic[0], zic[1:3], zic[3:5]
    return sos, zs


def _filtfilt_cascade(sos, zs, x):
    xf = rp.length_fixup(x)
    ext = np.concatenate([2 * xf[0] - xf[18:0:-1], xf, 2 * xf[-1] - xf[-2:-20:-1]])       # scipy odd_ext(x, 18)
    y1, _ = signal.sosfilt(sos, ext, zi=zs * ext[0])
    r = y1[::-1]
    y2, _ = signal.sosfilt(sos, r, zi=zs * r[0])
    return y2[::-1][18:-18]


def _df2t_longdouble(b, a, x, z0):
    """scipy's recurrence in 80-bit arithmetic: the exact response of (b, a, zi) for all purposes here."""
    ld = np.longdouble
    bq, aq, z, x = b.astype(ld), a.astype(ld), z0.astype(ld).copy(), x.astype(ld)
    y = np.empty(len(x), ld)
    for n in range(len(x)):
        yn = z[0] + bq[0] * x[n]
        for k in range(4):
            z[k] = (z[k + 1] + x[n] * bq[k + 1]) - yn * aq[k + 1]
        z[4] = x[n] * bq[5] - yn * aq[5]
        y[n] = yn
    return y


def test_sections_multiply_back_to_b_and_a():
    b, a, zi, sec, zic, m = _cascade()
    nb = np.polymul(np.polymul(sec[0, :2], sec[1, :3]), sec[2, :3])
    na = np.polymul(np.polymul([1.0, sec[0, 3]], [1.0, sec[1, 3], sec[1, 4]]), [1.0, sec[2, 3], sec[2, 4]])
    assert np.abs(nb - b).max() <= 4e-15 and np.abs(na - a).max() <= 4e-15
    assert sec[0, 2] == 0.0 and sec[0, 4] == 0.0                     # first-order section first
    # every section is stable and has its zeros next to z = 1 (a high-pass of its own)
    assert abs(sec[0, 3]) < 1 and all(abs(np.roots([1.0, s[3], s[4]])).max() < 1 for s in sec[1:])
    assert np.abs(m).max() < 100.0                                   # the DF2T realisation has 1e8 here


def test_cascade_matches_scipy_and_the_exact_response():
    b, a, zi, sec, zic, m = _cascade()
    sos, zs = _sos(sec, zic)
    rng = np.random.default_rng(11)
    t = np.arange(70000) / 16000.0
    base = 0.3 * np.sin(2 * np.pi * 110 * t) + 0.05 * rng.standard_normal(t.shape[0]) + 0.02
    for x in (base[:48000], base[:19], base[:257] + 0.5, np.zeros(3000), 0.7 * np.ones(5000), base[:256], base[:1000]):
        y = _filtfilt_cascade(sos, zs, x)
        assert np.abs(y - rp.highpass_filtfilt(rp.length_fixup(x))).max() <= 5e-7      # scipy's own fp64 wobble
    # forward pass against the 80-bit evaluation of scipy's recurrence from scipy's initial state
    x = base[:4000]
    y1, _ = signal.sosfilt(sos, x, zi=zs * x[0])
    exact = _df2t_longdouble(b, a, x, zi * x[0])
    assert np.abs(y1 - np.asarray(exact, np.float64)).max() <= 1e-9
    scipy_y, _ = signal.lfilter(b, a, x, zi=zi * x[0])
    assert np.abs(scipy_y - np.asarray(exact, np.float64)).max() > 10 * np.abs(y1 - np.asarray(exact, np.float64)).max()


def test_chunk_carry_is_plain_fp64():
    b, a, zi, sec, zic, m = _cascade(256)
    rng = np.random.default_rng(3)
    x = 0.2 * rng.standard_normal(256 * 6) + 0.1

    def run(xs, z):
        z = z.copy()
        for xn in xs:
            y0 = sec[0, 0] * xn + z[0]
            z[0] = sec[0, 1] * xn - sec[0, 3] * y0
            y1 = sec[1, 0] * y0 + z[1]
            z[1] = sec[1, 1] * y0 - sec[1, 3] * y1 + z[2]
            z[2] = sec[1, 2] * y0 - sec[1, 4] * y1
            y2 = sec[2, 0] * y1 + z[3]
            z[3] = sec[2, 1] * y1 - sec[2, 3] * y2 + z[4]
            z[4] = sec[2, 2] * y1 - sec[2, 4] * y2
        return z

    z = zic * x[0]
    for c in range(5):
        seg = x[256 * c:256 * (c + 1)]
        true_next = run(seg, z)
        carried = m @ z + run(seg, np.zeros(5))          # what filt_carry_kernel computes
        assert np.abs(carried - true_next).max() <= 1e-12
        z = true_next


def test_local_pass_taps_are_the_zero_state_finals():
    """ssfe_filt_cascade_taps: the dot products filt_dot_kernel takes equal the recurrence walked from a zero state."""
    b, a, zi, sec, zic, m = _cascade(256)
    lib = _lib.load()
    g = np.zeros((5, 256))
    assert lib.ssfe_filt_cascade_taps(np.ascontiguousarray(sec).ctypes.data, 256, g.ctypes.data) == 0
    assert lib.ssfe_filt_cascade_taps(None, 256, g.ctypes.data) == -1
    assert np.abs(g).max() < 2.0 and np.abs(g[:, -1]).max() > 1e-3          # well scaled: a plain fp64 sum is enough

    def run(xs):
        z = np.zeros(5, np.longdouble)
        s = sec.astype(np.longdouble)
        for xn in xs.astype(np.longdouble):
            y0 = s[0, 0] * xn + z[0]
            z[0] = s[0, 1] * xn - s[0, 3] * y0
            y1 = s[1, 0] * y0 + z[1]
            z[1] = s[1, 1] * y0 - s[1, 3] * y1 + z[2]
            z[2] = s[1, 2] * y0 - s[1, 4] * y1
            y2 = s[2, 0] * y1 + z[3]
            z[3] = s[2, 1] * y1 - s[2, 3] * y2 + z[4]
            z[4] = s[2, 2] * y1 - s[2, 4] * y2
        return np.asarray(z, np.float64)

    rng = np.random.default_rng(5)
    for x in (0.3 * rng.standard_normal(256) + 0.1, np.ones(256), np.eye(256)[0], np.eye(256)[255],
              np.round(3000 * rng.standard_normal(256)) / 32768.0):
        assert np.abs(g @ x - run(x)).max() <= 2e-13 * max(1.0, np.abs(x).max())
    # one chunk further: M z + taps . x is the true next state (what carry + final pass rely on)
    x = 0.2 * rng.standard_normal(512)
    z1 = g @ x[:256]
    assert np.abs(m @ z1 + g @ x[256:] - run(x)).max() <= 1e-12


def test_carry_powers_for_the_scan_over_runs():
    """ssfe_filt_cascade_powers: M^8, M^16, ... for the carry kernel's Kogge-Stone scan over runs of eight chunks."""
    b, a, zi, sec, zic, m = _cascade(256)
    lib = _lib.load()
    pw = np.zeros((5, 5, 5))
    assert lib.ssfe_filt_cascade_powers(np.ascontiguousarray(sec).ctypes.data, 256, 8, 5, pw.ctypes.data) == 0
    assert lib.ssfe_filt_cascade_powers(np.ascontiguousarray(sec).ctypes.data, 256, 8, 0, pw.ctypes.data) == -1
    ref = np.linalg.matrix_power(m.astype(np.longdouble), 8)
    for k in range(5):
        # (ref comes from the ROUNDED m: its 1e-16 rounding grows with the power)
        assert np.abs(pw[k] - np.asarray(ref, np.float64)).max() <= 1e-11 * max(1.0, np.abs(pw[k]).max())
        ref = ref @ ref
    # the scan's algebra: entry state of run l+1 = M^8 (entry of run l) + (run l walked from zero)
    rng = np.random.default_rng(9)
    s = rng.standard_normal((16, 5))
    z = rng.standard_normal(5)
    seq = [z]
    for c in range(16):
        seq.append(m @ seq[-1] + s[c])
    t0 = np.zeros(5)
    for c in range(8):
        t0 = m @ t0 + s[c]
    assert np.abs(pw[0] @ z + t0 - seq[8]).max() <= 1e-11 * max(1.0, np.abs(seq[8]).max())
