"""Restatement of ``librosa.filters.mel`` as called at /root/reference/make_spect_f0.py:15

    mel_basis = mel(16000, 1024, fmin=90, fmax=7600, n_mels=80).T

TEST INFRASTRUCTURE (oracle).  PARITY UNPINNED: librosa is a third-party dependency that is
not vendored in the reference and not installed here (README.md:28 lists it without a
version; the positional ``mel(sr, n_fft)`` call only works on librosa < 0.10).  This file
restates the published algorithm of that version range: Slaney mel scale (htk=False),
Slaney area normalisation (norm=1 / 'slaney'), float32 storage.
Cross-checks available in-container: torchaudio.functional.melscale_fbanks(...,'slaney','slaney') and
transformers.audio_utils.mel_filter_bank(norm='slaney', mel_scale='slaney') (tests/test_oracle.py).
Corroboration (not a pin): evaluated at other parameters the same code reproduces published outputs of the real
librosa.filters.mel - the first entries of Whisper's mel_filters.npz (sr 16000, n_fft 400, 80 filters) to all
eight printed digits, sign of the zero at [0, 0] included, and librosa's docstring example - quoted from memory
(tests/test_oracle.py::test_mel_restatement_reproduces_published_librosa_outputs).
"""
import numpy as np


def hz_to_mel(f):
    f = np.asanyarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp            # == 15
    logstep = np.log(6.4) / 27.0
    with np.errstate(divide="ignore", invalid="ignore"):
        log_t = f >= min_log_hz
        mels = np.where(log_t, min_log_mel + np.log(np.maximum(f, 1e-300) / min_log_hz) / logstep, mels)
    return mels


def mel_to_hz(m):
    m = np.asanyarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    freqs = f_sp * m
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    log_t = m >= min_log_mel
    return np.where(log_t, min_log_hz * np.exp(logstep * (m - min_log_mel)), freqs)


def mel_filterbank(sr=16000, n_fft=1024, n_mels=80, fmin=90.0, fmax=7600.0):
    """(n_mels, 1+n_fft//2) float32 triangular filterbank, Slaney-normalised."""
    n_bins = 1 + n_fft // 2
    weights = np.zeros((n_mels, n_bins), dtype=np.float32)
    fftfreqs = np.linspace(0.0, float(sr) / 2, n_bins, endpoint=True)
    mel_f = mel_to_hz(np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))      # stored as f32 first
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, np.newaxis]                                # f32 *= f64 -> f32
    return weights


def mel_basis_T():
    """The reference's module-level ``mel_basis`` (513, 80) float32 (make_spect_f0.py:15)."""
    return mel_filterbank().T
