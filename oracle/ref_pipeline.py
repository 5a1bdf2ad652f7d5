"""CPU restatement of the reference's feature front end (TEST INFRASTRUCTURE, not product).

Follows /root/reference/make_spect_f0.py:15-17,40-74 and /root/reference/utils.py:10-58 line by
line, in float64 like the reference.  scipy / numpy are the reference's own third-party
dependencies and ARE present in this image and on the GPU box, so ``filtfilt``, ``butter`` and
``RandomState`` are called directly (they are the reference's arithmetic, not a restatement).
``librosa.filters.mel`` -> oracle/mel_basis.py and ``pysptk.sptk.rapt`` -> oracle/rapt_ref.c are
restatements (PARITY UNPINNED for those two stages, see oracle/__init__.py).

Pinned against the reference's own ``utils.py`` by tests/golden/make_golden.py (run in the build
container where /root/reference exists) -> tests/golden/*.npz -> tests/test_oracle.py.
"""
import numpy as np
from numpy.random import RandomState
from scipy import signal
from scipy.signal import get_window

from .mel_basis import mel_basis_T

# make_spect_f0.py:16  min_level = exp(-100/20*ln 10) (== 1e-5 up to fp64 round-off)
MIN_LEVEL = np.exp(-100 / 20 * np.log(10))
UNVOICED = -1e10                     # make_spect_f0.py:65
GENDER_RANGE = {"M": (50, 250), "F": (100, 600)}   # make_spect_f0.py:40-45


def butter_highpass(cutoff=30, fs=16000, order=5):
    """utils.py:10-14."""
    nyq = 0.5 * fs
    normal_cutoff = cutoff / nyq
    b, a = signal.butter(order, normal_cutoff, btype="high", analog=False)
    return b, a


def length_fixup(x):
    """make_spect_f0.py:52-53: append one 1e-06 sample when L % 256 == 0."""
    x = np.asarray(x, dtype=np.float64)
    if x.shape[0] % 256 == 0:
        x = np.concatenate((x, np.array([1e-06])), axis=0)
    return x


def highpass_filtfilt(x, b=None, a=None):
    """make_spect_f0.py:54."""
    if b is None:
        b, a = butter_highpass(30, 16000, order=5)
    return signal.filtfilt(b, a, x)


def dither(y, prng):
    """make_spect_f0.py:55: consumes y.shape[0] doubles of the speaker stream."""
    return y * 0.96 + (prng.rand(y.shape[0]) - 0.5) * 1e-06


def pySTFT(x, fft_length=1024, hop_length=256):
    """utils.py:18-31 (1-D input only).  Returns (fft_length//2+1, T) f64."""
    x = np.pad(x, int(fft_length // 2), mode="reflect")
    noverlap = fft_length - hop_length
    n_frames = (x.shape[-1] - noverlap) // hop_length
    idx = np.arange(fft_length)[None, :] + hop_length * np.arange(n_frames)[:, None]
    frames = x[idx]
    fft_window = get_window("hann", fft_length, fftbins=True)
    result = np.fft.rfft(fft_window * frames, n=fft_length).T
    return np.abs(result)


def mel_db_normalize(D, mel_basis=None):
    """make_spect_f0.py:59-61.  D: (T,513) -> S (T,80) f64, NOT clipped."""
    if mel_basis is None:
        mel_basis = mel_basis_T()
    D_mel = np.dot(D, mel_basis)
    D_db = 20 * np.log10(np.maximum(MIN_LEVEL, D_mel)) - 16
    return (D_db + 100) / 100


def speaker_normalization(f0, index_nonzero, mean_f0, std_f0):
    """utils.py:35-42: voiced frames -> ((f0-mean)/std/4 clipped to [-1,1] + 1)/2, in float64;
    unvoiced frames keep their sentinel."""
    out = np.array(f0, dtype=np.float64)               # astype(float).copy()
    z = (out[index_nonzero] - mean_f0) / std_f0 / 4.0   # same operation order as :38
    z = np.clip(z, -1, 1)                               # :39
    out[index_nonzero] = (z + 1) / 2.0                  # :40
    return out


def quantize_f0_numpy(x, num_bins=256):
    """utils.py:46-58: uv = x<=0 -> bin 0; voiced -> round-half-even(x*255)+1; one-hot f32."""
    assert x.ndim == 1                                  # :48
    v = np.array(x, dtype=np.float64)                   # :49
    unvoiced = v <= 0                                   # :50
    v[unvoiced] = 0.0
    assert (v >= 0).all() and (v <= 1).all()            # :52
    bins = np.round(v * (num_bins - 1)) + 1             # :53-54 (np.round = half-to-even)
    bins[unvoiced] = 0.0                                # :55
    onehot = np.zeros((v.shape[0], num_bins + 1), dtype=np.float32)
    onehot[np.arange(v.shape[0]), bins.astype(np.int32)] = 1.0
    return onehot, bins.astype(np.int64)


def f0_stats(f0_rapt):
    """make_spect_f0.py:65-66 (float32 statistics over this utterance's voiced frames)."""
    index_nonzero = (f0_rapt != np.float32(UNVOICED))
    with np.errstate(all="ignore"):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            mean_f0, std_f0 = np.mean(f0_rapt[index_nonzero]), np.std(f0_rapt[index_nonzero])
    return index_nonzero, mean_f0, std_f0


def extract_utterance(x, gender, prng, mel_basis=None, ba=None, rapt_fn=None, want_stages=False):
    """One pass of the hot loop make_spect_f0.py:50-74 for an in-memory utterance.

    x       : float64 (L,) samples as ``sf.read`` returns them
    gender  : 'M' | 'F' (anything else raises ValueError like make_spect_f0.py:45)
    prng    : the speaker's RandomState (stream continues across calls)
    rapt_fn : callable(x_f32_scaled, fs, hop, lo, hi) -> f32 log-F0; default = C restatement
    Returns (S f32 (T,80), f0_norm f32 (T,)) - exactly what the script np.save's (:71-74).
    """
    if gender not in GENDER_RANGE:
        raise ValueError
    lo, hi = GENDER_RANGE[gender]
    if ba is None:
        ba = butter_highpass(30, 16000, order=5)
    if rapt_fn is None:
        from .rapt import rapt as rapt_fn
    x = length_fixup(x)
    y = highpass_filtfilt(x, *ba)
    wav = dither(y, prng)
    D = pySTFT(wav).T
    S = mel_db_normalize(D, mel_basis)
    f0_rapt = rapt_fn(wav.astype(np.float32) * 32768, 16000, 256, lo, hi)
    index_nonzero, mean_f0, std_f0 = f0_stats(f0_rapt)
    with np.errstate(all="ignore"):
        f0_norm = speaker_normalization(f0_rapt, index_nonzero, mean_f0, std_f0)
    assert len(S) == len(f0_rapt)
    out = (S.astype(np.float32), f0_norm.astype(np.float32))
    if want_stages:
        return out + (dict(y=y, wav=wav, D=D, S=S, f0_rapt=f0_rapt, mean=mean_f0, std=std_f0),)
    return out


def extract_speaker(spk, gender, utterances, **kw):
    """The inner loop (make_spect_f0.py:47-74) for one speaker directory ``p<int>``.

    utterances: list of float64 arrays in sorted(fileList) order.  One RandomState(int(spk[1:]))
    stream is shared by all of them, in order.
    """
    prng = RandomState(int(spk[1:]))
    return [extract_utterance(x, gender, prng, **kw) for x in utterances]
