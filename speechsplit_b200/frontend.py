"""FrontEnd: one context of libssfe.so on one GPU.

PyTorch is used here for device memory and streams only; all arithmetic runs in the
hand-written sm_100a kernels behind the C ABI (include/ssfe.h).  Method names follow the
reference stages they replace (make_spect_f0.py:50-74, utils.py:10-74).
"""
import ctypes
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np
import torch
from scipy import signal

from . import _lib as L
from .melbasis import reference_mel_basis

GENDER_RANGE = {"M": (50.0, 250.0), "F": (100.0, 600.0)}     # make_spect_f0.py:40-45


class SsfeError(RuntimeError):
    pass


@dataclass
class FrontEndConfig:
    """The hard-coded literals of make_spect_f0.py:15-17,55,60-61 in one place."""
    sample_rate: int = 16000
    n_fft: int = 1024
    hop: int = 256
    n_mels: int = 80
    fmin: float = 90.0
    fmax: float = 7600.0
    cutoff_hz: float = 30.0
    order: int = 5
    wav_scale: float = 0.96
    dither_scale: float = 1e-06
    ref_db: float = 16.0
    min_level: float = float(np.exp(-100 / 20 * np.log(10)))
    filtfilt_mode: int = 0        # 1 = sequential validation mode
    mel_basis: Optional[np.ndarray] = field(default=None, repr=False)


def butter_highpass(cutoff, fs, order=5):
    """reference utils.py:10-14 (host-side constant, the reference's own scipy call)."""
    nyq = 0.5 * fs
    normal_cutoff = cutoff / nyq
    b, a = signal.butter(order, normal_cutoff, btype="high", analog=False)
    return b, a


def _raise(code, msg):
    if code == L.SSFE_ERR_RANGE:
        raise AssertionError(msg)                 # utils.py:52 / :68
    if code in (L.SSFE_ERR_TOO_SHORT, L.SSFE_ERR_GENDER):
        raise ValueError(msg)                     # scipy / pysptk ValueError, make_spect_f0.py:45
    if code == L.SSFE_ERR_NOMEM:
        raise MemoryError(msg)
    raise SsfeError("libssfe error %d: %s" % (code, msg))


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _ptr(a, typ):
    return a.ctypes.data_as(typ)


_DT = {torch.float32: L.F32, torch.float64: L.F64, torch.int16: L.I16}
_NP_DT = {np.dtype(np.float32): L.F32, np.dtype(np.float64): L.F64, np.dtype(np.int16): L.I16}


def _wav_dtype(table, dt):
    if dt not in table:
        raise TypeError("samples must be float32, float64 or int16 PCM, not %s" % (dt,))
    return table[dt]


def _check_ragged(off, numel, what="sample_offsets"):
    """The C ABI takes plain pointers: offsets that leave the buffer would be an out-of-bounds device
    read, so they are refused here (the reference indexes numpy arrays, which raise by themselves)."""
    if off.ndim != 1 or off.shape[0] < 1:
        raise ValueError("%s must be a 1-D array of n + 1 offsets" % what)
    if off[0] < 0 or np.any(np.diff(off) < 0):
        raise ValueError("%s must be non-negative and non-decreasing" % what)
    if int(off[-1]) > int(numel):
        raise ValueError("%s end at %d but the buffer holds %d elements" % (what, int(off[-1]), int(numel)))


def _check_out(name, a, shape, dtype, device=None):
    """A caller-supplied output buffer must be exactly what the kernels write: shape, dtype, dense."""
    if isinstance(a, torch.Tensor):
        ok = tuple(a.shape) == tuple(shape) and a.dtype == dtype and a.is_contiguous() and \
            (device is None or a.device == device) and (device is not None or a.device.type == "cpu")
    else:
        npdt = {torch.float32: np.float32, torch.float64: np.float64, torch.int64: np.int64}[dtype]
        ok = device is None and isinstance(a, np.ndarray) and a.shape == tuple(shape) and a.dtype == npdt and \
            a.flags.c_contiguous and a.flags.writeable
    if not ok:
        raise ValueError("output buffer %r must be a dense %s array of shape %s%s" %
                         (name, str(dtype).replace("torch.", ""), tuple(shape),
                          "" if device is None else " on %s" % (device,)))


class FrontEnd:
    def __init__(self, device: int = 0, config: Optional[FrontEndConfig] = None):
        if not torch.cuda.is_available():
            raise RuntimeError("speechsplit_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = L.load()
        self.config = config or FrontEndConfig()
        self.device = torch.device("cuda", device)
        c = self.config
        b, a = butter_highpass(c.cutoff_hz, c.sample_rate, c.order)
        zi = signal.lfilter_zi(b, a)
        mel = c.mel_basis if c.mel_basis is not None else reference_mel_basis()
        self._mel = np.ascontiguousarray(mel, dtype=np.float32)
        assert self._mel.shape == (c.n_fft // 2 + 1, c.n_mels)
        cfg = L.Config()
        cfg.sample_rate, cfg.n_fft, cfg.hop, cfg.n_mels = c.sample_rate, c.n_fft, c.hop, c.n_mels
        cfg.b[:] = list(b)
        cfg.a[:] = list(a)
        cfg.zi[:] = list(zi)
        cfg.mel_basis = _ptr(self._mel, L.c_f32p)
        cfg.min_level, cfg.ref_db = c.min_level, c.ref_db
        cfg.wav_scale, cfg.dither_scale = c.wav_scale, c.dither_scale
        cfg.filtfilt_mode = c.filtfilt_mode
        self.b, self.a, self.zi = b, a, zi
        h = L.vp()
        with torch.cuda.device(self.device):
            torch.cuda.init()
            rc = self.lib.ssfe_create(ctypes.byref(h), device, ctypes.byref(cfg))
        if rc != 0:
            _raise(rc, self.lib.ssfe_last_error(None).decode())
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self.lib.ssfe_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- helpers -----------------------------------------------------------------------------
    def _check(self, rc):
        if rc != 0:
            _raise(rc, self.lib.ssfe_last_error(self._h).decode())

    def _bind_stream(self):
        """Enqueue on torch's current stream so results are ordered with the caller's tensors.
        torch's default stream has handle 0, which the C ABI reads as "use the context's own
        stream"; pass CUDA's explicit legacy-default handle (cudaStreamLegacy == 0x1) instead."""
        h = torch.cuda.current_stream(self.device).cuda_stream
        self._check(self.lib.ssfe_set_stream(self._h, L.vp(h if h else 1)))

    def _dev(self, t, dtype=None):
        if isinstance(t, np.ndarray):
            t = torch.from_numpy(np.ascontiguousarray(t))
        t = t.to(self.device, non_blocking=False)
        if dtype is not None and t.dtype != dtype:
            t = t.to(dtype)
        return t.contiguous()

    @property
    def launch_count(self):
        return int(self.lib.ssfe_launch_count(self._h))

    STAGES = ("rand", "filtfilt", "edges", "stft_mel", "rapt_decimate", "rapt_cand", "rapt_stat", "rapt_dp", "f0_post")

    def enable_timing(self, on=True):
        self._check(self.lib.ssfe_enable_timing(self._h, 1 if on else 0))

    def stage_ms(self):
        """Milliseconds per stage of the most recent timed extract() (CUDA events on the launch stream)."""
        buf = (ctypes.c_float * 16)()
        n = self.lib.ssfe_stage_ms(self._h, buf, 16)
        if n < 0:
            self._check(n)
        return {k: float(buf[i]) for i, k in enumerate(self.STAGES[:n])}

    def synchronize(self):
        self._check(self.lib.ssfe_synchronize(self._h))

    @staticmethod
    def plan(sample_offsets):
        """(fixed_offsets, frame_offsets) of a ragged batch (make_spect_f0.py:52-53, :69)."""
        so = _i64(sample_offsets)
        n = so.shape[0] - 1
        fix = np.empty(n + 1, np.int64)
        fr = np.empty(n + 1, np.int64)
        rc = L.load().ssfe_plan_offsets(_ptr(so, L.c_i64p), n, _ptr(fix, L.c_i64p), _ptr(fr, L.c_i64p))
        if rc != 0:
            raise ValueError("bad sample offsets")
        return fix, fr

    # ---- stages ------------------------------------------------------------------------------
    def filtfilt(self, x, sample_offsets):
        """make_spect_f0.py:52-54 on a ragged batch -> float64 tensor [fixed total]."""
        x = self._dev(x)
        so = _i64(sample_offsets)
        _check_ragged(so, x.numel())
        n = so.shape[0] - 1
        fix, _ = self.plan(so)
        y = torch.empty(int(fix[-1]), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            self._bind_stream()
            self._check(self.lib.ssfe_filtfilt(self._h, L.vp(x.data_ptr()), _wav_dtype(_DT, x.dtype),
                                               _ptr(so, L.c_i64p), n, L.vp(y.data_ptr())))
        return y, fix

    def rand(self, seeds, skips, counts):
        """RandomState(seed).rand(): counts[i] doubles after skipping skips[i] (make_spect_f0.py:47,55)."""
        seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
        skips = np.ascontiguousarray(skips, dtype=np.uint64)
        off = np.concatenate([[0], np.cumsum(_i64(counts))]).astype(np.int64)
        n = seeds.shape[0]
        u = torch.empty(int(off[-1]), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            self._bind_stream()
            self._check(self.lib.ssfe_rand(self._h, _ptr(seeds, L.c_u32p), _ptr(skips, L.c_u64p),
                                           _ptr(off, L.c_i64p), n, L.vp(u.data_ptr())))
        return u, off

    def _stft(self, fn, width, wav, offsets):
        wav = self._dev(wav, torch.float32)
        off = _i64(offsets)
        _check_ragged(off, wav.numel(), "offsets")
        n = off.shape[0] - 1
        frames = (np.diff(off) + self.config.hop) // self.config.hop
        out = torch.empty((int(frames.sum()), width), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            self._bind_stream()
            self._check(fn(self._h, L.vp(wav.data_ptr()), _ptr(off, L.c_i64p), n, L.vp(out.data_ptr())))
        return out, np.concatenate([[0], np.cumsum(frames)]).astype(np.int64)

    def stft_mag(self, wav, offsets):
        """utils.pySTFT (utils.py:18-31), frame-major: float32 [frames, 513]."""
        return self._stft(self.lib.ssfe_stft_mag, self.config.n_fft // 2 + 1, wav, offsets)

    def stft_mel_db(self, wav, offsets):
        """make_spect_f0.py:58-61 fused: float32 [frames, 80] (unclipped S)."""
        return self._stft(self.lib.ssfe_stft_mel_db, self.config.n_mels, wav, offsets)

    def rapt(self, wav, offsets, f0_lo, f0_hi):
        """make_spect_f0.py:64: log-F0 float32 [sum ceil(L/256)], unvoiced -1e10.  wav is NOT pre-scaled."""
        wav = self._dev(wav, torch.float32)
        off = _i64(offsets)
        _check_ragged(off, wav.numel(), "offsets")
        n = off.shape[0] - 1
        lo = np.ascontiguousarray(f0_lo, dtype=np.float32)
        hi = np.ascontiguousarray(f0_hi, dtype=np.float32)
        frames = -(-np.diff(off) // self.config.hop)
        out = torch.empty(int(frames.sum()), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            self._bind_stream()
            self._check(self.lib.ssfe_rapt(self._h, L.vp(wav.data_ptr()), _ptr(off, L.c_i64p), n,
                                           _ptr(lo, L.c_f32p), _ptr(hi, L.c_f32p), L.vp(out.data_ptr())))
        return out, np.concatenate([[0], np.cumsum(frames)]).astype(np.int64)

    def rapt_dump(self, max_frames):
        """Per-frame records of the most recent RAPT run (diagnostics for the parity tests)."""
        m = int(max_frames)
        d = dict(ncands=np.zeros(m, np.uint8), locs=np.zeros((m, 20), np.int16), mpvals=np.zeros((m, 20), np.float32),
                 f0cand=np.zeros((m, 20), np.float32), stat=np.zeros(m, np.float32), rms_ratio=np.zeros(m, np.float32))
        n = self.lib.ssfe_rapt_dump(self._h, m, *[L.vp(d[k].ctypes.data) for k in
                                                  ("ncands", "locs", "mpvals", "f0cand", "stat", "rms_ratio")])
        if n < 0:
            self._check(int(n))
        return {k: v[:n] for k, v in d.items()}

    def f0_normalize(self, f0, frame_offsets):
        """make_spect_f0.py:65-67 per utterance -> (f0_norm f32, stats f32 [n,2])."""
        f0 = self._dev(f0, torch.float32)
        fo = _i64(frame_offsets)
        n = fo.shape[0] - 1
        out = torch.empty_like(f0)
        stats = torch.empty((n, 2), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            self._bind_stream()
            self._check(self.lib.ssfe_f0_normalize(self._h, L.vp(f0.data_ptr()), _ptr(fo, L.c_i64p), n,
                                                   L.vp(out.data_ptr()), L.vp(stats.data_ptr())))
        return out, stats

    def speaker_normalization(self, f0, index_nonzero, mean_f0, std_f0):
        """utils.speaker_normalization (utils.py:35-42) -> float64 tensor."""
        f0 = self._dev(f0)
        if f0.dtype not in (torch.float32, torch.float64):
            f0 = f0.double()
        nz = self._dev(index_nonzero).to(torch.uint8).contiguous()
        out = torch.empty(f0.shape, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            self._bind_stream()
            self._check(self.lib.ssfe_speaker_normalization(self._h, L.vp(f0.data_ptr()), _DT[f0.dtype],
                                                            L.vp(nz.data_ptr()), float(mean_f0), float(std_f0),
                                                            f0.numel(), L.vp(out.data_ptr())))
        return out

    def quantize_f0(self, x, num_bins=256, check_range=True, want_onehot=True):
        """utils.quantize_f0_numpy / _torch (utils.py:46-74) over a flat tensor."""
        x = self._dev(x)
        if x.dtype not in (torch.float32, torch.float64):
            x = x.double()
        cnt = x.numel()
        onehot = torch.empty((cnt, num_bins + 1), dtype=torch.float32, device=self.device) if want_onehot else None
        bins = torch.empty(cnt, dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            self._bind_stream()
            self._check(self.lib.ssfe_quantize_f0(self._h, L.vp(x.data_ptr()), _DT[x.dtype], cnt, num_bins,
                                                  L.vp(onehot.data_ptr() if want_onehot else 0),
                                                  L.vp(bins.data_ptr()), 1 if check_range else 0))
        return onehot, bins

    # ---- the whole hot loop --------------------------------------------------------------------
    def _batch(self, sample_offsets, f0_lo, f0_hi, spk_seed, dither_skip):
        so = _i64(sample_offsets)
        n = so.shape[0] - 1
        keep = dict(so=so, lo=np.ascontiguousarray(f0_lo, dtype=np.float32),
                    hi=np.ascontiguousarray(f0_hi, dtype=np.float32),
                    seed=np.ascontiguousarray(spk_seed, dtype=np.uint32),
                    skip=np.ascontiguousarray(dither_skip, dtype=np.uint64))
        for k in ("lo", "hi", "seed", "skip"):
            if keep[k].shape[0] != n:
                raise ValueError("batch array %s has %d entries for %d utterances" % (k, keep[k].shape[0], n))
        b = L.Batch()
        b.n_utts = n
        b.sample_offsets = _ptr(so, L.c_i64p)
        b.f0_lo = _ptr(keep["lo"], L.c_f32p)
        b.f0_hi = _ptr(keep["hi"], L.c_f32p)
        b.spk_seed = _ptr(keep["seed"], L.c_u32p)
        b.dither_skip = _ptr(keep["skip"], L.c_u64p)
        return b, keep

    def extract(self, x, sample_offsets, f0_lo, f0_hi, spk_seed, dither_skip, want=("mel", "f0_norm"), out=None):
        """make_spect_f0.py:50-74 for a ragged batch resident in HBM.

        x: 1-D device tensor (float32 / float64 / int16 PCM) holding the concatenated utterances.
        want: any of mel, f0_norm, f0_raw, onehot, bins, wav, wav64.  Returns dict of device tensors
        (+ 'fixed_offsets', 'frame_offsets' numpy arrays).
        """
        x = self._dev(x)
        b, keep = self._batch(sample_offsets, f0_lo, f0_hi, spk_seed, dither_skip)
        _check_ragged(keep["so"], x.numel())
        fix, fr = self.plan(keep["so"])
        T, S = int(fr[-1]), int(fix[-1])
        shapes = dict(mel=((T, self.config.n_mels), torch.float32), f0_norm=((T,), torch.float32),
                      f0_raw=((T,), torch.float32), onehot=((T, 257), torch.float32), bins=((T,), torch.int64),
                      wav=((S,), torch.float32), wav64=((S,), torch.float64))
        want = set(want) | {"mel", "f0_norm"}
        res = {}
        o = L.Outputs()
        for k in want:
            shp, dt = shapes[k]
            if out is not None and k in out:
                t = out[k]
                _check_out(k, t, shp, dt, self.device)
            else:
                t = torch.empty(shp, dtype=dt, device=self.device)
            res[k] = t
            setattr(o, k, t.data_ptr())
        with torch.cuda.device(self.device):
            self._bind_stream()
            self._check(self.lib.ssfe_extract(self._h, ctypes.byref(b), L.vp(x.data_ptr()),
                                              _wav_dtype(_DT, x.dtype), ctypes.byref(o)))
        res["fixed_offsets"], res["frame_offsets"] = fix, fr
        return res

    def extract_host(self, x_host, sample_offsets, f0_lo, f0_hi, spk_seed, dither_skip, want_bins=True, out=None):
        """Same with HOST buffers (numpy arrays or pinned torch CPU tensors) in and out."""
        if isinstance(x_host, torch.Tensor):
            xt = x_host
            if xt.device.type != "cpu" or not xt.is_contiguous():
                raise ValueError("extract_host takes a dense CPU tensor (use extract for device tensors)")
            ptr, dt, numel = xt.data_ptr(), _wav_dtype(_DT, xt.dtype), xt.numel()
        else:
            x_host = np.ascontiguousarray(x_host)
            ptr, dt, numel = x_host.ctypes.data, _wav_dtype(_NP_DT, x_host.dtype), x_host.size
        b, keep = self._batch(sample_offsets, f0_lo, f0_hi, spk_seed, dither_skip)
        _check_ragged(keep["so"], numel)
        fix, fr = self.plan(keep["so"])
        T = int(fr[-1])
        if out is None:
            out = dict(mel=np.empty((T, self.config.n_mels), np.float32), f0_norm=np.empty(T, np.float32))
            if want_bins:
                out["bins"] = np.empty(T, np.int64)
        else:
            _check_out("mel", out["mel"], (T, self.config.n_mels), torch.float32)
            _check_out("f0_norm", out["f0_norm"], (T,), torch.float32)
            if out.get("bins") is not None:
                _check_out("bins", out["bins"], (T,), torch.int64)

        def hp(a):
            if a is None:
                return L.vp(0)
            return L.vp(a.data_ptr() if isinstance(a, torch.Tensor) else a.ctypes.data)

        with torch.cuda.device(self.device):
            self._bind_stream()
            self._check(self.lib.ssfe_extract_host(self._h, ctypes.byref(b), L.vp(ptr), dt, hp(out["mel"]),
                                                   hp(out["f0_norm"]), hp(out.get("bins"))))
        out["fixed_offsets"], out["frame_offsets"] = fix, fr
        return out

    def collate(self, mel, f0_norm, frame_offsets, utt, left, len_crop, max_len_pad=192, want_onehot=True):
        """data_loader.py:101-128 on the GPU (+ the solver.py:162 quantisation of the padded F0)."""
        mel = self._dev(mel, torch.float32)
        f0_norm = self._dev(f0_norm, torch.float32)
        fo = _i64(frame_offsets)
        utt = np.ascontiguousarray(utt, dtype=np.int32)
        left = np.ascontiguousarray(left, dtype=np.int32)
        len_crop = np.ascontiguousarray(len_crop, dtype=np.int32)
        n = utt.shape[0]
        if left.shape[0] != n or len_crop.shape[0] != n:
            raise ValueError("collate: utt, left and len_crop must have one entry per item")
        _check_ragged(fo, f0_norm.numel(), "frame_offsets")
        if mel.numel() != f0_norm.numel() * self.config.n_mels:
            raise ValueError("collate: mel must hold %d values per F0 frame" % self.config.n_mels)
        if n and (utt.min() < 0 or utt.max() >= fo.shape[0] - 1):
            raise ValueError("collate: utterance index out of range (%d utterances)" % (fo.shape[0] - 1))
        melsp = torch.empty((n, max_len_pad, self.config.n_mels), dtype=torch.float32, device=self.device)
        pitch = torch.empty((n, max_len_pad, 1), dtype=torch.float32, device=self.device)
        onehot = torch.empty((n, max_len_pad, 257), dtype=torch.float32, device=self.device) if want_onehot else None
        bins = torch.empty((n, max_len_pad), dtype=torch.int64, device=self.device) if want_onehot else None
        with torch.cuda.device(self.device):
            self._bind_stream()
            self._check(self.lib.ssfe_collate(self._h, L.vp(mel.data_ptr()), L.vp(f0_norm.data_ptr()),
                                              _ptr(fo, L.c_i64p), n, _ptr(utt, L.c_i32p), _ptr(left, L.c_i32p),
                                              _ptr(len_crop, L.c_i32p), max_len_pad, L.vp(melsp.data_ptr()),
                                              L.vp(pitch.data_ptr()), L.vp(onehot.data_ptr() if want_onehot else 0),
                                              L.vp(bins.data_ptr() if want_onehot else 0)))
        return melsp, pitch, onehot, bins


_default = {}


def default_frontend(device: Optional[int] = None) -> FrontEnd:
    """Process-wide context per GPU, created on first use."""
    if device is None:
        device = torch.cuda.current_device() if torch.cuda.is_available() else 0
    if device not in _default:
        _default[device] = FrontEnd(device)
    return _default[device]
