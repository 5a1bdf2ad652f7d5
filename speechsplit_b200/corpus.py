"""Synthetic 16 kHz corpora shaped like the reference's ``assets/wavs/<spk>/*.wav`` tree.

The reference ships no audio; tests and bench.py therefore use speech-like
synthetic utterances: a harmonic source whose F0 wanders inside the
speaker's range, shaped by three formant resonances, alternating with noise bursts and
silences, over a 1e-4 white-noise floor, peak 0.1-0.5, quantised to 16-bit PCM exactly as a
.wav file would hold it (``sf.read`` then yields int16/32768 as float64, make_spect_f0.py:50).

Speakers are directories ``p<int>``; the integer seeds the per-speaker dither stream
(make_spect_f0.py:47) and ``spk2gen`` maps them to 'M'/'F' (make_spect_f0.py:19,40-45).
"""
from dataclasses import dataclass
from typing import List, Sequence

import numpy as np
import torch

FS = 16000
F0_RANGE = {"M": (80.0, 180.0), "F": (150.0, 350.0)}


@dataclass(frozen=True)
class UttMeta:
    spk: str          # 'p226'
    gender: str       # 'M' | 'F'
    index: int        # position in sorted(fileList) of that speaker
    length: int       # samples on disk (before the make_spect_f0.py:52-53 fix-up)
    seed: int         # synthesis seed

    @property
    def spk_id(self) -> int:
        return int(self.spk[1:])


def speaker_gender(spk_id: int) -> str:
    """Fixed synthetic spk2gen table: odd ids female, even ids male."""
    return "F" if spk_id % 2 else "M"


def make_manifest(n_speakers=109, utts_per_speaker=400, first_id=225, mean_s=3.0, std_s=0.8,
                  min_s=1.0, max_s=8.0, seed=0, fixed_len=None) -> List[UttMeta]:
    """VCTK-shaped manifest, speakers sorted, utterances in file order.  About 1/256 of the
    lengths land on a multiple of 256 by chance; every 61st utterance is forced onto one so the
    append path (make_spect_f0.py:52-53) is always exercised."""
    rng = np.random.Generator(np.random.PCG64(seed))
    metas = []
    for s in range(n_speakers):
        sid = first_id + s
        g = speaker_gender(sid)
        for u in range(utts_per_speaker):
            if fixed_len is not None:
                L = int(fixed_len)
            else:
                dur = float(np.clip(rng.normal(mean_s, std_s), min_s, max_s))
                L = int(round(dur * FS))
                if (s * utts_per_speaker + u) % 61 == 0:
                    L -= L % 256
            metas.append(UttMeta("p%d" % sid, g, u, L, 1000 * sid + u))
    return metas


def _control_tracks(meta: UttMeta, n_ctl: int):
    """100 Hz control tracks: f0 (Hz), voiced gain, noise gain."""
    rng = np.random.Generator(np.random.PCG64(meta.seed))
    lo, hi = F0_RANGE[meta.gender]
    f0 = np.empty(n_ctl, np.float32)
    vg = np.zeros(n_ctl, np.float32)
    ng = np.zeros(n_ctl, np.float32)
    t = 0
    cur = rng.uniform(lo, hi)
    state = 0 if rng.random() < 0.7 else 2
    while t < n_ctl:
        if state == 0:      # voiced 150-600 ms
            n = int(rng.integers(15, 61))
            steps = rng.normal(0.0, 0.012, n).cumsum()
            seg = np.clip(cur * np.exp(steps), lo, hi)
            cur = float(seg[-1])
            env = np.minimum(1.0, np.minimum(np.arange(1, n + 1), np.arange(n, 0, -1)) / 3.0)
            e = min(n, n_ctl - t)
            f0[t:t + e] = seg[:e]
            vg[t:t + e] = (rng.uniform(0.5, 1.0) * env)[:e]
        elif state == 1:    # noise burst 40-150 ms
            n = int(rng.integers(4, 16))
            e = min(n, n_ctl - t)
            f0[t:t + e] = cur
            ng[t:t + e] = rng.uniform(0.05, 0.3)
        else:               # silence 50-200 ms
            n = int(rng.integers(5, 21))
            e = min(n, n_ctl - t)
            f0[t:t + e] = cur
        t += n
        state = int(rng.choice(3, p=[0.6, 0.2, 0.2]))
    formants = np.array([rng.uniform(400, 900), rng.uniform(1100, 2300), rng.uniform(2500, 3500)], np.float32)
    peak = np.float32(rng.uniform(0.1, 0.5))
    tilt = np.float32(rng.uniform(1.0, 2.0))
    return f0, vg, ng, formants, peak, tilt


def control_tracks(meta: UttMeta):
    """Picklable per-utterance job (bench.py maps it over a process pool before CUDA starts)."""
    return _control_tracks(meta, meta.length // 160 + 2)


@torch.no_grad()
def synth_batch(metas: Sequence[UttMeta], device="cpu", n_harm=14, tracks=None) -> List[torch.Tensor]:
    """Synthesise utterances; returns one int16 tensor (length,) per meta, on ``device``.
    ``tracks``: optional precomputed ``control_tracks(meta)`` results, same order."""
    if len(metas) == 0:
        return []
    dev = torch.device(device)
    B = len(metas)
    Lmax = max(m.length for m in metas)
    n_ctl = Lmax // 160 + 2
    f0c = np.zeros((B, n_ctl), np.float32)
    vgc = np.zeros((B, n_ctl), np.float32)
    ngc = np.zeros((B, n_ctl), np.float32)
    fm = np.zeros((B, 3), np.float32)
    pk = np.zeros(B, np.float32)
    tl = np.zeros(B, np.float32)
    for i, m in enumerate(metas):
        a, b_, c, fm[i], pk[i], tl[i] = tracks[i] if tracks is not None else control_tracks(m)
        k = a.shape[0]
        f0c[i, :k], vgc[i, :k], ngc[i, :k] = a, b_, c
        f0c[i, k:] = a[-1]

    def up(a):
        ta = torch.from_numpy(a).to(dev)[:, None, :]
        return torch.nn.functional.interpolate(ta, size=(n_ctl - 1) * 160 + 1, mode="linear",
                                               align_corners=True)[:, 0, :Lmax]

    f0 = up(f0c)
    vg = up(vgc)
    ng = up(ngc)
    phase = torch.cumsum(f0.double() * (2.0 * np.pi / FS), dim=1)
    phase = torch.remainder(phase, 2.0 * np.pi).float()
    fm_t = torch.from_numpy(fm).to(dev)
    tilt = torch.from_numpy(tl).to(dev)[:, None]
    harm = torch.zeros_like(f0)
    for h in range(1, n_harm + 1):
        fh = f0 * h
        gain = torch.zeros_like(f0)
        for k, bw in enumerate((120.0, 180.0, 260.0)):
            gain = gain + 1.0 / (1.0 + ((fh - fm_t[:, k:k + 1]) / bw) ** 2)
        gain = (gain + 0.05) * (float(h) ** (-tilt)) * (fh < 0.45 * FS)
        harm = harm + gain * torch.sin(h * phase)
    gen = torch.Generator(device=dev)
    gen.manual_seed(int(metas[0].seed) * 7919 + B)
    noise = torch.randn(f0.shape, generator=gen, device=dev)
    sig = vg * harm + ng * noise
    sig = sig / sig.abs().amax(dim=1, keepdim=True).clamp_min(1e-6) * torch.from_numpy(pk).to(dev)[:, None]
    sig = sig + 1e-4 * torch.randn(f0.shape, generator=gen, device=dev)
    pcm = torch.clamp(torch.round(sig * 32768.0), -32768, 32767).to(torch.int16)
    return [pcm[i, :m.length].contiguous() for i, m in enumerate(metas)]


def pcm_to_float64(pcm) -> np.ndarray:
    """What ``sf.read`` returns for 16-bit PCM (make_spect_f0.py:50): int16 / 32768 as float64."""
    if isinstance(pcm, torch.Tensor):
        pcm = pcm.cpu().numpy()
    return pcm.astype(np.float64) / 32768.0
