"""CPU tests: the oracle (oracle/) against the reference-generated golden vectors
(tests/golden/make_golden.py) and against known answers of SURVEY.md 8(c)."""
import os

import numpy as np
import pytest
from numpy.random import RandomState

from oracle import ref_pipeline as rp
from oracle.mel_basis import mel_filterbank
from oracle.rapt import rapt
from speechsplit_b200.corpus import UttMeta, pcm_to_float64


@pytest.fixture(scope="module")
def kat(golden_dir):
    return np.load(os.path.join(golden_dir, "utils_kat.npz"))


def test_butter_coefficients(kat):
    b, a = rp.butter_highpass(30, 16000, order=5)
    assert np.array_equal(b, kat["b"]) and np.array_equal(a, kat["a"])
    assert abs(a.sum() - 2.2266566e-10) < 1e-15     # SURVEY.md hard part 2


def test_mt19937_kat(kat):
    assert np.array_equal(RandomState(226).rand(8), kat["rand_226"])


def test_pystft_matches_reference(kat):
    D = rp.pySTFT(kat["stft_x"])
    assert D.shape == (513, 20) and D.dtype == np.float64
    assert np.allclose(D[:, [0, 1, 7, 19]], kat["stft_D_cols"], rtol=0, atol=1e-12)
    assert abs(D.sum() - kat["stft_D_sum"]) < 1e-8


def test_frame_counts():
    for L, T in ((48000, 188), (48128, 189), (32768, 129), (960000, 3751)):
        x = rp.length_fixup(np.zeros(L))
        assert (x.shape[0] + 256) // 256 == T
        assert -(-x.shape[0] // 256) == T            # == ceil(L'/256), RAPT's frame count


def test_speaker_normalization_matches_reference(kat):
    f0 = kat["sn_f0"]
    out = rp.speaker_normalization(f0, f0 != -1e10, kat["sn_mean"], kat["sn_std"])
    assert out.dtype == np.float64 and np.array_equal(out, kat["sn_out"])
    idx, mean, std = rp.f0_stats(f0)
    assert mean == kat["sn_mean"] and std == kat["sn_std"] and mean.dtype == np.float32


def test_quantize_matches_reference(kat):
    enc, idx = rp.quantize_f0_numpy(kat["q_in"])
    assert enc.dtype == np.float32 and idx.dtype == np.int64 and enc.shape == (len(idx), 257)
    assert np.array_equal(idx, kat["q_idx"])
    assert np.array_equal(enc.argmax(1), kat["q_enc_argmax"]) and np.array_equal(enc.sum(1), kat["q_enc_sum"])
    assert list(idx[:6]) == [0, 0, 1, 129, 256, 65]   # SURVEY.md 3.4 (half-to-even tie)
    with pytest.raises(AssertionError):
        rp.quantize_f0_numpy(np.array([0.5, 1.5]))
    with pytest.raises(AssertionError):
        rp.quantize_f0_numpy(np.zeros((2, 2)))


def test_mel_basis_structure():
    w = mel_filterbank()
    assert w.shape == (80, 513) and w.dtype == np.float32
    assert int((w != 0).sum()) == 941
    cols = np.nonzero(w.any(0))[0]
    assert cols[0] == 6 and cols[-1] == 486 and int((w != 0).sum(0).max()) == 2


def test_mel_basis_vs_torchaudio():
    ta = pytest.importorskip("torchaudio")
    t = ta.functional.melscale_fbanks(513, 90.0, 7600.0, 80, 16000, "slaney", "slaney").numpy().T
    assert np.abs(mel_filterbank() - t).max() < 2e-7


def test_mel_basis_vs_transformers():
    """Second independent restatement of librosa.filters.mel (transformers.audio_utils documents its
    slaney/slaney filter bank as librosa's): same support, at most one f32 ulp apart."""
    au = pytest.importorskip("transformers.audio_utils")
    t = au.mel_filter_bank(513, 80, 90.0, 7600.0, 16000, norm="slaney", mel_scale="slaney").T.astype(np.float32)
    w = mel_filterbank()
    assert np.array_equal(w != 0, t != 0)
    assert np.abs(w.view(np.int32) - t.view(np.int32)).max() <= 1


def test_mel_restatement_reproduces_published_librosa_outputs():
    """Two outputs of the real ``librosa.filters.mel`` that are public, evaluated with the restatement at THEIR
    parameters (the function is the same one, only sr / n_fft / fmin / fmax / n_mels differ from the reference's call):

    * OpenAI Whisper ships ``mel_filters.npz`` = ``librosa.filters.mel(sr=16000, n_fft=400, n_mels=80)`` (so its
      audio.py says); its widely printed first entries are ``-0., 0.02486259`` in filter 0 and ``0.00199082,
      0.02287177`` at the start of filter 1;
    * librosa's own docstring example ``librosa.filters.mel(sr=22050, n_fft=2048)`` prints ``[0., 0.016, ...]``.

    PROVENANCE: neither file is in this image - the numbers are quoted from memory of those public artefacts, which
    makes this a corroboration of the restated formulae (mel scale break, area normalisation, f32 storage, even the
    negative zero at [0, 0]), not the pin that tests/test_oracle_thirdparty.py becomes on a box with librosa."""
    w = mel_filterbank(16000, 400, 80, 0.0, 8000.0)
    assert w.shape == (80, 201) and w.dtype == np.float32
    assert w[0, 0] == 0.0 and np.signbit(w[0, 0])
    assert abs(float(w[0, 1]) - 0.02486259) < 5e-9 and w[0, 2] == 0.0
    assert abs(float(w[1, 1]) - 0.00199082) < 5e-9 and abs(float(w[1, 2]) - 0.02287177) < 5e-9
    d = mel_filterbank(22050, 2048, 128, 0.0, 11025.0)
    assert d.shape == (128, 1025) and d[0, 0] == 0.0 and round(float(d[0, 1]), 3) == 0.016


@pytest.mark.parametrize("name", ["pipeline_p226.npz", "pipeline_p225.npz"])
def test_pipeline_matches_reference(golden_dir, name):
    """make_spect_f0.py:47-74 through the oracle == through the reference's own functions."""
    g = np.load(os.path.join(golden_dir, name))
    spk, gender, n = str(g["spk"]), str(g["gender"]), int(g["n"])
    utts = [pcm_to_float64(g["pcm%d" % k]) for k in range(n)]
    prng = RandomState(int(spk[1:]))
    for k, x in enumerate(utts):
        S, f0n, st = rp.extract_utterance(x, gender, prng, want_stages=True)
        assert np.array_equal(S, g["S%d" % k])
        assert np.array_equal(st["f0_rapt"], g["f0_rapt%d" % k])
        assert np.array_equal(f0n, g["f0_norm%d" % k], equal_nan=True)
        assert np.array_equal(rp.quantize_f0_numpy(f0n)[1], g["bins%d" % k])
        if k == 0:
            assert np.array_equal(st["y"], g["y0"]) and np.array_equal(st["wav"], g["wav0"])
    # the second p226 file has L % 256 == 0 -> the append path of :52-53
    if name == "pipeline_p226.npz":
        assert g["pcm1"].shape[0] % 256 == 0 and g["S1"].shape[0] == g["pcm1"].shape[0] // 256 + 1


def test_unknown_gender_raises():
    with pytest.raises(ValueError):
        rp.extract_utterance(np.zeros(4000), "X", RandomState(1))


def test_rapt_contract():
    """Output length ceil(L/hop), sentinel -1e10, too-short input raises (pysptk ValueError)."""
    rng = np.random.default_rng(0)
    t = np.arange(16000) / 16000.0
    x = (0.3 * np.sin(2 * np.pi * 120 * t) + 0.1 * np.sin(2 * np.pi * 240 * t) + 1e-3 * rng.standard_normal(16000))
    f0 = rapt(x.astype(np.float32) * 32768, 16000, 256, 50, 250)
    assert f0.dtype == np.float32 and f0.shape[0] == 63
    v = f0 != np.float32(-1e10)
    assert v.sum() > 50 and np.all(np.abs(np.exp(f0[v]) - 120.0) < 1.0)
    assert np.all(f0[-2:] == np.float32(-1e10))         # tail frames RAPT cannot analyse
    sil = rapt((1e-6 * (rng.random(16000) - 0.5)).astype(np.float32) * 32768, 16000, 256, 50, 250)
    assert np.all(sil == np.float32(-1e10))
    with pytest.raises(ValueError):
        rapt(np.zeros(600, np.float32), 16000, 256, 50, 250)


def test_rapt_debug_consistency():
    rng = np.random.default_rng(1)
    t = np.arange(40000) / 16000.0
    f = 180 + 40 * np.sin(2 * np.pi * 0.7 * t)
    x = 0.2 * np.sin(2 * np.pi * np.cumsum(f) / 16000.0) * (np.sin(2 * np.pi * 1.3 * t) > -0.3)
    x = x + 1e-4 * rng.standard_normal(x.shape[0])
    f0, d = rapt(x.astype(np.float32) * 32768, 16000, 256, 100, 600, debug=True)
    T = d["n_frames"]
    assert T <= f0.shape[0] and T >= f0.shape[0] - 4
    nc = d["ncands"][:T]
    assert nc.min() >= 1 and nc.max() <= 20
    for i in range(T):                      # last candidate of every frame is the unvoiced one
        assert d["locs"][i, nc[i] - 1] == -1
        assert np.all(d["locs"][i, :nc[i] - 1] >= 27) and np.all(d["locs"][i, :nc[i] - 1] <= 160)


def test_interp_lnr_oracle_matches_reference_module(golden_dir):
    """oracle/interp_lnr.py vs the output of the reference's own model.InterpLnr (training mode,
    model.py:380-436) for the same captured random draws: bit-identical."""
    from oracle.interp_lnr import interp_lnr
    z = np.load(os.path.join(golden_dir, "interp_lnr.npz"))
    for k in range(int(z["n"])):
        y = interp_lnr(z["x%d" % k], z["len_seq%d" % k], z["scales%d" % k], z["len_seg%d" % k])
        assert y.dtype == np.float32 and np.array_equal(y, z["y%d" % k]), k


def test_collator_matches_reference(golden_dir):
    """oracle.collate_ref == the reference's own data_loader.MyCollator (data_loader.py:96-128) on the same
    items and the same numpy seed, and == the train.pkl of the reference's make_metadata.py."""
    from oracle import collate_ref
    g = np.load(os.path.join(golden_dir, "collate.npz"))
    items = [(g["S%d" % k], g["emb%d" % k], g["f0%d" % k]) for k in range(int(g["n"]))]
    np.random.seed(int(g["seed"]))
    melsp, spk_emb, pitch, len_org = collate_ref.collate([items[i] for i in g["order"]], 64, 128, 192)
    for got, name in ((melsp, "melsp"), (spk_emb, "spk_emb"), (pitch, "pitch"), (len_org, "len_org")):
        assert got.dtype == g[name].dtype and np.array_equal(got, g[name]), name
    assert melsp.shape == (8, 192, 80) and pitch.shape == (8, 192, 1) and len_org.dtype == np.int64
    tree = {}
    for e in g["meta_tree"]:
        spk, f = str(e).split("/")
        tree.setdefault(spk, []).append(f)
    meta = collate_ref.metadata(tree)
    assert [m[0] for m in meta] == [str(x) for x in g["meta_speakers"]]
    assert np.array_equal(np.stack([m[1] for m in meta]), g["meta_emb"])
    assert ["|".join(m[2:]) for m in meta] == [str(x) for x in g["meta_files"]]


def test_demo_asset_quantisation_matches_reference(golden_dir):
    """The two real normalised-F0 tracks of the reference's assets/demo.pkl (outputs of the original pipeline
    with the real pysptk) through the oracle's quantize_f0_numpy == through the reference's (demo.ipynb:36-50)."""
    g = np.load(os.path.join(golden_dir, "demo_kat.npz"))
    for k in range(int(g["n"])):
        f0 = g["f0%d" % k]
        f0_pad = np.pad(f0, (0, 192 - len(f0)), "constant", constant_values=(0, 0))
        enc, idx = rp.quantize_f0_numpy(f0_pad)
        assert np.array_equal(idx, g["idx%d" % k])
        assert np.array_equal(enc.argmax(1), g["enc_argmax%d" % k]) and np.array_equal(enc.sum(1), g["enc_sum%d" % k])
        assert (idx[:len(f0)] > 0).sum() > 30 and np.all(idx[len(f0):] == 0)      # voiced frames exist; padding is bin 0


def test_torch_eager_restatement_matches_reference_vectors(golden_dir, kat):
    """oracle/torch_eager_ref.py (the reference's own GPU-side eager ops, bench.py's opponent for configs[4]) on
    the CPU: InterpLnr vs the output of the reference's module for the captured draws, quantize_f0_torch vs the
    vectors the reference's utils.py produced."""
    import torch
    from oracle.torch_eager_ref import interp_lnr_eager, quantize_f0_eager
    z = np.load(os.path.join(golden_dir, "interp_lnr.npz"))
    for k in range(int(z["n"])):
        x = torch.from_numpy(z["x%d" % k])
        y = interp_lnr_eager(x, torch.from_numpy(z["len_seq%d" % k]), max_len_pad=z["y%d" % k].shape[1],
                             draws=(torch.from_numpy(z["scales%d" % k]), torch.from_numpy(z["len_seg%d" % k])))
        assert np.array_equal(y.numpy(), z["y%d" % k]), k
    enc, idx = quantize_f0_eager(torch.from_numpy(kat["qt_in"]))
    assert enc.dtype == torch.float32 and idx.dtype == torch.int64
    assert np.array_equal(idx.numpy(), kat["qt_idx"]) and np.array_equal(enc.argmax(-1).numpy(), kat["qt_enc_argmax"])
