"""Generates tests/golden/*.npz by running the REFERENCE's own code (``/root/reference/utils.py``
and the arithmetic of ``make_spect_f0.py:52-61`` with the reference's functions) in the build
container.  /root/reference does not exist on the GPU box, so the vectors are committed.

    python tests/golden/make_golden.py

What is reference-produced (pins the oracle and the CUDA path):
  utils_kat.npz      butter_highpass, pySTFT, speaker_normalization, quantize_f0_numpy,
                     quantize_f0_torch, pad_seq_to_2 outputs of the reference's utils.py
  pipeline_*.npz     y / wav / D_mel-dB-normalised S of make_spect_f0.py:52-61 executed with the
                     reference's butter_highpass + pySTFT + scipy filtfilt + numpy RandomState
  interp_lnr.npz     inputs, captured random draws and output of the reference's model.InterpLnr
                     (model.py:380-436) in training mode
  collate.npz        inputs, numpy seed and output batch of the reference's data_loader.MyCollator
                     (data_loader.py:96-128) and the train.pkl of the reference's make_metadata.py
  demo_kat.npz       the two real (VCTK p226 / p231) normalised-F0 tracks shipped in the reference's
                     assets/demo.pkl - outputs of the ORIGINAL pipeline with the real pysptk / librosa - and
                     what the reference's utils.pad_seq_to_2 + quantize_f0_numpy make of them (demo.ipynb:36-50)
What is NOT reference-produced (librosa / pysptk are absent):
  the mel basis (oracle/mel_basis.py restatement) used inside pipeline_*.npz, and the
  ``f0_rapt`` arrays (oracle/rapt_ref.c restatement) - stored as regression vectors and
  flagged ``rapt_pinned=False``.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

# utils.py:5 imports librosa.filters.mel and never uses it; stub the import only.
_lf = types.ModuleType("librosa.filters")
_lf.mel = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("librosa is not installed"))
_l = types.ModuleType("librosa")
_l.filters = _lf
sys.modules["librosa"] = _l
sys.modules["librosa.filters"] = _lf
sys.path.insert(0, "/root/reference")
import utils as ref_utils  # noqa: E402  (the reference's own module)
from numpy.random import RandomState  # noqa: E402
from scipy import signal  # noqa: E402

from oracle.mel_basis import mel_basis_T  # noqa: E402
from oracle.rapt import rapt  # noqa: E402
from speechsplit_b200.corpus import UttMeta, synth_batch, pcm_to_float64  # noqa: E402


def utils_kat():
    rng = np.random.Generator(np.random.PCG64(7))
    b, a = ref_utils.butter_highpass(30, 16000, order=5)
    x = rng.standard_normal(5000) * 0.1
    D = ref_utils.pySTFT(x)
    f0 = np.where(rng.random(300) < 0.3, -1e10, rng.normal(5.0, 0.3, 300)).astype(np.float32)
    nz = f0 != -1e10
    mean, std = np.mean(f0[nz]), np.std(f0[nz])
    f0n = ref_utils.speaker_normalization(f0, nz, mean, std)
    q_in = np.concatenate([np.array([-1e10, 0.0, 1e-9, 0.5, 1.0, 0.25]),
                           (np.arange(0, 255) + 0.5) / 255.0, rng.random(200),
                           f0n.astype(np.float32)])
    q_enc, q_idx = ref_utils.quantize_f0_numpy(q_in)
    qt_in = torch.from_numpy(np.where(rng.random((3, 192)) < 0.3, -1e10, rng.random((3, 192))).astype(np.float32))
    qt_enc, qt_idx = ref_utils.quantize_f0_torch(qt_in)
    pad_in = rng.random((1, 135, 80)).astype(np.float32)
    pad_out, len_pad = ref_utils.pad_seq_to_2(pad_in, 192)
    np.savez_compressed(os.path.join(HERE, "utils_kat.npz"),
                        b=b, a=a, stft_x=x, stft_D=D.astype(np.float32), stft_D_sum=D.sum(),
                        stft_D_cols=D[:, [0, 1, 7, 19]],
                        sn_f0=f0, sn_mean=mean, sn_std=std, sn_out=f0n,
                        q_in=q_in, q_idx=q_idx, q_enc_argmax=q_enc.argmax(1).astype(np.int16),
                        q_enc_sum=q_enc.sum(1), qt_in=qt_in.numpy(), qt_idx=qt_idx.numpy(),
                        qt_enc_argmax=qt_enc.argmax(-1).numpy().astype(np.int16),
                        pad_in_sum=pad_in.sum(), pad_out_shape=np.array(pad_out.shape), len_pad=len_pad,
                        rand_226=RandomState(226).rand(8))


def pipeline(name, metas):
    """make_spect_f0.py:47-74 for one speaker with the reference's own helper functions."""
    mel_basis = mel_basis_T()
    min_level = np.exp(-100 / 20 * np.log(10))
    b, a = ref_utils.butter_highpass(30, 16000, order=5)
    pcm = synth_batch(metas)
    prng = RandomState(metas[0].spk_id)
    lo, hi = (50, 250) if metas[0].gender == "M" else (100, 600)
    out = {"rapt_pinned": False, "gender": metas[0].gender, "spk": metas[0].spk}
    for k, (m, p) in enumerate(zip(metas, pcm)):
        x = pcm_to_float64(p)
        if x.shape[0] % 256 == 0:
            x = np.concatenate((x, np.array([1e-06])), axis=0)
        y = signal.filtfilt(b, a, x)
        wav = y * 0.96 + (prng.rand(y.shape[0]) - 0.5) * 1e-06
        D = ref_utils.pySTFT(wav).T
        D_mel = np.dot(D, mel_basis)
        D_db = 20 * np.log10(np.maximum(min_level, D_mel)) - 16
        S = (D_db + 100) / 100
        f0_rapt = rapt(wav.astype(np.float32) * 32768, 16000, 256, lo, hi)
        nz = f0_rapt != -1e10
        mean_f0, std_f0 = np.mean(f0_rapt[nz]), np.std(f0_rapt[nz])
        f0_norm = ref_utils.speaker_normalization(f0_rapt, nz, mean_f0, std_f0)
        assert len(S) == len(f0_rapt)
        enc, idx = ref_utils.quantize_f0_numpy(f0_norm.astype(np.float32))
        out["pcm%d" % k] = p.numpy()
        out["S%d" % k] = S.astype(np.float32)
        out["f0_rapt%d" % k] = f0_rapt
        out["f0_norm%d" % k] = f0_norm.astype(np.float32)
        out["bins%d" % k] = idx.astype(np.int16)
        out["mean_std%d" % k] = np.array([mean_f0, std_f0], np.float32)
        if k == 0:
            out["y0"] = y
            out["wav0"] = wav
    out["n"] = len(metas)
    np.savez_compressed(os.path.join(HERE, name), **out)


def interp_lnr():
    """InterpLnr.forward of the reference's model.py (training mode), run on the CPU with a seeded
    generator.  The two random tensors it draws (model.py:392-393, :401-404) are captured by making the
    same two calls after the same seed - the module's forward then consumes the identical stream."""
    import model as ref_model  # the reference's own module (needs only torch, numpy and utils)

    class HP:
        max_len_seq, max_len_pad, min_len_seg, max_len_seg = 128, 192, 19, 32

    mod = ref_model.InterpLnr(HP)
    mod.train()
    out = {}
    for k, (seed, B, C) in enumerate(((0, 4, 81), (7, 16, 8), (11, 3, 5))):
        g = torch.Generator().manual_seed(1000 + seed)
        x = torch.rand((B, HP.max_len_pad, C), generator=g, dtype=torch.float32)
        len_seq = torch.randint(HP.max_len_seq // 2, HP.max_len_seq + 1, (B,), generator=g)
        if k == 2:
            len_seq[0] = 2                                   # degenerate: a single interpolation interval
            len_seq[1] = HP.max_len_pad                      # longer than any draw can cover
        torch.manual_seed(seed)
        scales = torch.rand(B * mod.max_num_seg) + 0.5
        len_seg = torch.randint(low=HP.min_len_seg, high=HP.max_len_seg, size=(B * mod.max_num_seg, 1))
        torch.manual_seed(seed)
        y = mod(x, len_seq)
        out["x%d" % k], out["len_seq%d" % k] = x.numpy(), len_seq.numpy()
        out["scales%d" % k], out["len_seg%d" % k] = scales.numpy(), len_seg.numpy().reshape(-1)
        out["y%d" % k] = y.numpy()
    out["n"] = 3
    np.savez_compressed(os.path.join(HERE, "interp_lnr.npz"), **out)


def collator():
    """The reference's own data_loader.MyCollator and make_metadata.py, run unmodified.  Two things are
    supplied from outside, neither touches their arithmetic: the name ``pdb`` (data_loader.py:106 calls
    ``pdb.set_trace()`` but the module never imports pdb - a no-op object is put into the module's globals),
    and a working directory that holds ``assets/spmel`` (make_metadata.py is a script with relative paths)."""
    import pickle
    import runpy
    import tempfile

    import data_loader as ref_dl  # the reference's own module
    ref_dl.pdb = types.SimpleNamespace(set_trace=lambda: None)

    class HP:
        min_len_seq, max_len_seq, max_len_pad = 64, 128, 192

    rng = np.random.Generator(np.random.PCG64(99))
    Ts = [135, 200, 129, 400, 131]
    out = {"n": len(Ts), "seed": 314, "order": np.array([3, 0, 4, 1, 2, 3, 3, 1], np.int64)}
    items = []
    for k, T in enumerate(Ts):
        S = (rng.random((T, 80)) * 1.3 - 0.15).astype(np.float32)
        f0 = rng.random(T).astype(np.float32)
        f0[rng.random(T) < 0.3] = -1e10
        emb = np.zeros(82, np.float32)
        emb[1 if k == 1 else 7] = 1.0
        out["S%d" % k], out["f0%d" % k], out["emb%d" % k] = S, f0, emb
        items.append((S, emb, f0))
    np.random.seed(int(out["seed"]))
    melsp, spk_emb, pitch, len_org = ref_dl.MyCollator(HP)([items[i] for i in out["order"]])
    out["melsp"], out["spk_emb"], out["pitch"], out["len_org"] = (melsp.numpy(), spk_emb.numpy(), pitch.numpy(),
                                                                   len_org.numpy())
    # make_metadata.py in a scratch tree: speakers and files created in non-sorted order
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        tree = {"p300": ["p300_010.npy", "p300_002.npy"], "p226": ["p226_005.npy"], "p225": ["p225_003.npy", "p225_001.npy", "p225_002.npy"]}
        for spk, files in tree.items():
            os.makedirs(os.path.join(tmp, "assets", "spmel", spk))
            for f in files:
                np.save(os.path.join(tmp, "assets", "spmel", spk, f), np.zeros((1, 80), np.float32))
        os.chdir(tmp)
        try:
            runpy.run_path("/root/reference/make_metadata.py")
        finally:
            os.chdir(cwd)
        with open(os.path.join(tmp, "assets", "spmel", "train.pkl"), "rb") as fh:
            meta = pickle.load(fh)
    out["meta_tree"] = np.array(["%s/%s" % (s, f) for s, fs in tree.items() for f in fs])
    out["meta_speakers"] = np.array([m[0] for m in meta])
    out["meta_emb"] = np.stack([m[1] for m in meta])
    out["meta_files"] = np.array(["|".join(m[2:]) for m in meta])
    np.savez_compressed(os.path.join(HERE, "collate.npz"), **out)


def demo_kat():
    """demo.ipynb:36-50 on the reference's own assets/demo.pkl with the reference's own utils functions."""
    import pickle
    with open("/root/reference/assets/demo.pkl", "rb") as fh:
        demo = pickle.load(fh)
    out = {"n": len(demo)}
    for k, sbmt in enumerate(demo):
        mel, f0, length = sbmt[2][0], sbmt[2][1], sbmt[2][2]
        out["spk%d" % k], out["mel%d" % k], out["f0%d" % k], out["len%d" % k] = sbmt[0], mel, f0, length
        mel_pad, len_pad = ref_utils.pad_seq_to_2(mel[np.newaxis, :, :], 192)
        f0_pad = np.pad(f0, (0, 192 - len(f0)), "constant", constant_values=(0, 0))
        enc, idx = ref_utils.quantize_f0_numpy(f0_pad)
        out["mel_pad%d" % k], out["len_pad%d" % k] = mel_pad, len_pad
        out["enc_argmax%d" % k], out["enc_sum%d" % k], out["idx%d" % k] = enc.argmax(1), enc.sum(1), idx
    np.savez_compressed(os.path.join(HERE, "demo_kat.npz"), **out)


if __name__ == "__main__":
    if sys.argv[1:] == ["demo"]:
        demo_kat()
        sys.exit(0)
    if sys.argv[1:] == ["collate"]:          # only the loader vectors (the others are unchanged)
        collator()
        sys.exit(0)
    utils_kat()
    interp_lnr()
    collator()
    demo_kat()
    # cfg1: one 3.000 s male utterance, speaker p226 (+ two more files of the same speaker so the
    # dither stream continuity and the L % 256 == 0 append path are pinned)
    pipeline("pipeline_p226.npz", [UttMeta("p226", "M", 0, 48000, 226000),
                                   UttMeta("p226", "M", 1, 32768, 226001),
                                   UttMeta("p226", "M", 2, 20011, 226002)])
    pipeline("pipeline_p225.npz", [UttMeta("p225", "F", 0, 40000, 225000),
                                   UttMeta("p225", "F", 1, 25600, 225001)])
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
