"""Drop-in mirror of the reference's ``utils.py`` for the hot path (same names, argument meaning,
return dtypes/shapes and exceptions), computed by the sm_100a kernels of libssfe.so.

    reference                          here
    utils.butter_highpass   :10-14     host-side scipy call, unchanged (a 12-number constant)
    utils.pySTFT            :18-31     ssfe_stft_mag        (fp32 FFT on the GPU, returned as f64)
    utils.speaker_normalization :35-42 ssfe_speaker_normalization
    utils.quantize_f0_numpy :46-58     ssfe_quantize_f0
    utils.quantize_f0_torch :62-74     ssfe_quantize_f0 (result stays on x.device)
    get_mask_from_lengths / pad_seq_to_2 :78-88  trivial host helpers kept for the consumers

No CPU fallback: importing works anywhere, calling needs a B200.
"""
import numpy as np
import torch

from .frontend import butter_highpass, default_frontend  # noqa: F401  (butter_highpass re-exported)


def pySTFT(x, fft_length=1024, hop_length=256):
    """|STFT| of a 1-D signal: float64 (fft_length//2+1, T), T = (len(x)+hop)//hop.

    Only 1-D input is meaningful in the reference too."""
    if fft_length != 1024 or hop_length != 256:
        raise ValueError("the sm_100a kernel is specialised for fft_length=1024, hop_length=256")
    x = np.asarray(x)
    if x.ndim != 1:
        raise ValueError("pySTFT: only 1-D input is supported")
    fe = default_frontend()
    mag, _ = fe.stft_mag(torch.from_numpy(x.astype(np.float32)), [0, x.shape[0]])
    return mag.t().double().cpu().numpy()


def speaker_normalization(f0, index_nonzero, mean_f0, std_f0):
    """f0 is logf0.  Voiced entries -> ((f0-mean)/std/4 clipped to [-1,1] + 1)/2; float64 copy."""
    f0 = np.asarray(f0)
    idx = np.asarray(index_nonzero)
    if idx.dtype != np.bool_:            # integer index arrays: expand to a mask
        m = np.zeros(f0.shape, dtype=bool)
        m[idx] = True
        idx = m
    fe = default_frontend()
    src = f0 if f0.dtype in (np.float32, np.float64) else f0.astype(np.float64)
    out = fe.speaker_normalization(torch.from_numpy(np.ascontiguousarray(src)),
                                   torch.from_numpy(np.ascontiguousarray(idx)), float(mean_f0), float(std_f0))
    return out.cpu().numpy()


def quantize_f0_numpy(x, num_bins=256):
    """x is (normalised) logf0; returns (float32 (T, num_bins+1) one-hot, int64 (T,))."""
    assert x.ndim == 1
    x = np.asarray(x)
    src = x if x.dtype in (np.float32, np.float64) else x.astype(np.float64)
    fe = default_frontend()
    enc, idx = fe.quantize_f0(torch.from_numpy(np.ascontiguousarray(src)), num_bins, check_range=True)
    return enc.cpu().numpy(), idx.cpu().numpy()


def quantize_f0_torch(x, num_bins=256):
    """x: float tensor (B, T) -> (float32 (B, T, num_bins+1), int64 (B, T)) on x.device."""
    B = x.size(0)
    fe = default_frontend(x.device.index if x.is_cuda else None)
    enc, idx = fe.quantize_f0(x.reshape(-1), num_bins, check_range=True)
    enc, idx = enc.view(B, -1, num_bins + 1), idx.view(B, -1)
    if not x.is_cuda:
        enc, idx = enc.cpu(), idx.cpu()
    return enc, idx


def get_mask_from_lengths(lengths, max_len):
    ids = torch.arange(0, max_len, device=lengths.device)
    return (ids >= lengths.unsqueeze(1)).bool()


def pad_seq_to_2(x, len_out=128):
    len_pad = len_out - x.shape[1]
    assert len_pad >= 0
    return np.pad(x, ((0, 0), (0, len_pad), (0, 0)), "constant"), len_pad
