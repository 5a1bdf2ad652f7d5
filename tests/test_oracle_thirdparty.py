"""The two stages whose parity is UNPINNED: `librosa.filters.mel` (make_spect_f0.py:15) and
`pysptk.sptk.rapt` (make_spect_f0.py:64).  Neither package is in the image, in /opt/wheelhouse or under
baseline/_ref, so these tests SKIP today; the first box that has them turns them into the pin:
the oracle's restatements are compared with the real packages on the golden PCM."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_REF = os.path.join(ROOT, "baseline", "_ref")
if os.path.isdir(_REF) and _REF not in sys.path:      # an offline install of the reference's dependencies, if any
    sys.path.append(_REF)


def _golden_wavs(golden_dir):
    """(gender, float64 wav after filtfilt + dither) of the golden utterances (tests/golden/pipeline_*.npz)."""
    out = []
    for name in ("pipeline_p225.npz", "pipeline_p226.npz"):
        g = np.load(os.path.join(golden_dir, name))
        out.append((str(g["gender"]), np.asarray(g["wav0"], np.float64)))
    return out


def test_mel_basis_vs_librosa():
    librosa = pytest.importorskip("librosa")
    from oracle.mel_basis import mel_basis_T
    try:
        ref = librosa.filters.mel(16000, 1024, fmin=90, fmax=7600, n_mels=80).T          # make_spect_f0.py:15 (librosa < 0.10)
    except TypeError:
        ref = librosa.filters.mel(sr=16000, n_fft=1024, fmin=90, fmax=7600, n_mels=80).T
    ours = mel_basis_T()
    assert ours.shape == ref.shape == (513, 80)
    assert np.array_equal(ours != 0, ref != 0), "support differs"
    assert np.abs(ours.astype(np.float64) - ref.astype(np.float64)).max() <= 1.2e-7        # one f32 ulp of the largest weight


def test_rapt_vs_pysptk(golden_dir):
    pysptk = pytest.importorskip("pysptk")
    from oracle.rapt import rapt
    for gender, wav in _golden_wavs(golden_dir):
        lo, hi = (50, 250) if gender == "M" else (100, 600)
        x = wav.astype(np.float32) * 32768                                               # make_spect_f0.py:64
        ref = pysptk.sptk.rapt(x, 16000, 256, min=lo, max=hi, otype=2)
        ours = rapt(x, 16000, 256, lo, hi)
        assert ours.shape == ref.shape
        vr, vo = ref != np.float32(-1e10), ours != np.float32(-1e10)
        assert (vr == vo).mean() >= 0.999, "voicing flags differ on %.3f %% of frames" % (100 * (vr != vo).mean())
        both = vr & vo
        cents = 1731.234 * np.abs(ref[both].astype(np.float64) - ours[both].astype(np.float64))
        assert cents.max() <= 1.0, "worst F0 deviation %.3f cent" % cents.max()


def test_whole_loop_vs_real_packages(golden_dir):
    """make_spect_f0.py:52-67 with the real librosa + pysptk against oracle.ref_pipeline on the golden PCM."""
    pytest.importorskip("librosa")
    pysptk = pytest.importorskip("pysptk")
    from numpy.random import RandomState
    from oracle import ref_pipeline as rp
    g = np.load(os.path.join(golden_dir, "pipeline_p226.npz"))
    x = g["pcm0"].astype(np.float64) / 32768.0
    fn = lambda w, fs, hop, lo, hi: pysptk.sptk.rapt(w, fs, hop, min=lo, max=hi, otype=2)
    S0, f0 = rp.extract_utterance(x, str(g["gender"]), RandomState(226))
    S1, f1 = rp.extract_utterance(x, str(g["gender"]), RandomState(226), rapt_fn=fn)
    assert np.array_equal(S0, S1)
    assert (rp.quantize_f0_numpy(f0)[1] == rp.quantize_f0_numpy(f1)[1]).mean() >= 0.999
