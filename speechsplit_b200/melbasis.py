"""Host-side constant of the front end: the (513, 80) mel basis of reference make_spect_f0.py:15

    mel_basis = librosa.filters.mel(16000, 1024, fmin=90, fmax=7600, n_mels=80).T

If librosa is importable it is called exactly like the reference does; otherwise the same
matrix is built here (Slaney mel scale, Slaney area normalisation, float32 storage - the
defaults of every librosa release that accepts the reference's positional call).
"""
import numpy as np

_F_SP = 200.0 / 3
_MIN_LOG_HZ = 1000.0
_MIN_LOG_MEL = _MIN_LOG_HZ / _F_SP
_LOGSTEP = np.log(6.4) / 27.0


def _hz_to_mel(f):
    f = np.atleast_1d(np.asarray(f, dtype=np.float64))
    out = f / _F_SP
    hi = f >= _MIN_LOG_HZ
    out[hi] = _MIN_LOG_MEL + np.log(f[hi] / _MIN_LOG_HZ) / _LOGSTEP
    return out


def _mel_to_hz(m):
    m = np.atleast_1d(np.asarray(m, dtype=np.float64))
    out = _F_SP * m
    hi = m >= _MIN_LOG_MEL
    out[hi] = _MIN_LOG_HZ * np.exp(_LOGSTEP * (m[hi] - _MIN_LOG_MEL))
    return out


def slaney_mel_filterbank(sr, n_fft, n_mels, fmin, fmax):
    n_bins = 1 + n_fft // 2
    fft_hz = np.linspace(0.0, sr / 2.0, n_bins)
    edges = _mel_to_hz(np.linspace(_hz_to_mel(fmin)[0], _hz_to_mel(fmax)[0], n_mels + 2))
    width = np.diff(edges)
    dist = edges[:, None] - fft_hz[None, :]                   # (n_mels+2, n_bins)
    rising = -dist[:-2] / width[:-1, None]
    falling = dist[2:] / width[1:, None]
    tri = np.maximum(0, np.minimum(rising, falling)).astype(np.float32)   # librosa fills a float32 array
    tri *= (2.0 / (edges[2:] - edges[:-2]))[:, None]                       # then scales it in place
    return tri


def reference_mel_basis():
    """(513, 80) float32, C-contiguous."""
    try:
        from librosa.filters import mel  # the reference's own call, if the package exists
        try:
            m = mel(16000, 1024, fmin=90, fmax=7600, n_mels=80)
        except TypeError:   # librosa >= 0.10 made the arguments keyword-only
            m = mel(sr=16000, n_fft=1024, fmin=90, fmax=7600, n_mels=80)
    except ImportError:
        m = slaney_mel_filterbank(16000, 1024, 80, 90.0, 7600.0)
    return np.ascontiguousarray(m.T.astype(np.float32))
