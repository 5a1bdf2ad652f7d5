"""GPU parity tests, stage by stage, through the C ABI (speechsplit_b200.FrontEnd -> libssfe.so)
against the CPU oracle (oracle/) and the reference-generated golden vectors.  Bars:
  integer / index work (MT19937 stream, bins, one-hot)  : bit-exact
  float32 statistics (mean / std)                       : bit-exact (numpy pairwise order reproduced)
  filtfilt (fp64)                                       : <= 1e-6 abs vs scipy (scipy's own sequential
                                                          fp64 round-off for this filter is ~2e-7, see
                                                          DESIGN.md "IIR conditioning")
  |STFT| (fp32 FFT)                                     : <= 1e-5 * frame peak magnitude
  normalised mel                                        : <= 1e-4 abs  (BASELINE.json north_star)
"""
import os

import numpy as np
import pytest
import torch
from numpy.random import RandomState

from oracle import ref_pipeline as rp
from speechsplit_b200.corpus import make_manifest, pcm_to_float64, synth_batch

pytestmark = pytest.mark.gpu


def _ragged(arrs, dtype):
    off = np.concatenate([[0], np.cumsum([len(a) for a in arrs])]).astype(np.int64)
    return np.concatenate(arrs).astype(dtype), off


# ---- a2: the dither stream --------------------------------------------------------------------
def test_rand_bit_exact_single_stream(fe):
    u, off = fe.rand([226], [0], [100000])
    assert np.array_equal(u.cpu().numpy(), RandomState(226).rand(100000))


def test_rand_stream_continuity_and_many_speakers(fe):
    """Per-speaker streams continue across files (make_spect_f0.py:47-48,55); requests arrive in any
    order, with gaps, from many speakers at once."""
    rng = np.random.default_rng(3)
    seeds, skips, counts, expect = [], [], [], []
    for spk in (225, 226, 300, 4000000000):
        ref = RandomState(spk).rand(60000)
        pos = 0
        for _ in range(6):
            gap = int(rng.integers(0, 700)) if rng.random() < 0.4 else 0
            cnt = int(rng.integers(1, 9000))
            seeds.append(spk), skips.append(pos + gap), counts.append(cnt)
            expect.append(ref[pos + gap:pos + gap + cnt])
            pos += gap + cnt
    perm = rng.permutation(len(seeds))
    u, off = fe.rand(np.array(seeds, np.uint32)[perm], np.array(skips)[perm], np.array(counts)[perm])
    u = u.cpu().numpy()
    for k, j in enumerate(perm):
        assert np.array_equal(u[off[k]:off[k + 1]], expect[j]), "request %d" % j


def test_rand_block_boundaries(fe):
    ref = RandomState(7).rand(2000)
    for skip, cnt in ((0, 1), (311, 2), (312, 312), (623, 1), (0, 624), (937, 5)):
        u, _ = fe.rand([7], [skip], [cnt])
        assert np.array_equal(u.cpu().numpy(), ref[skip:skip + cnt])


def _stream_slice(seed, skip, cnt):
    rs = RandomState(seed)
    left = skip
    while left > 0:
        step = min(left, 20_000_000)
        rs.rand(step)
        left -= step
    return rs.rand(cnt)


def test_rand_segment_jumps(fe):
    """The generator cuts every stream into segments of 4096 state blocks (1 277 952 doubles) and
    jumps to their start states (mt19937.cu / mt_jump.cpp): requests that cross segment edges, start
    deep inside a stream, share a segment, or leave whole segments untouched must still be the
    bits numpy produces."""
    seg = 4096 * 312
    ref = RandomState(226).rand(3 * seg + 5000)
    cases = [(seg - 1000, 5000), (seg, 312), (2 * seg - 1, 2), (0, 2 * seg + 17), (3 * seg - 7, 3000)]
    for skip, cnt in cases:
        u, _ = fe.rand([226], [skip], [cnt])
        assert np.array_equal(u.cpu().numpy(), ref[skip:skip + cnt]), (skip, cnt)
    # several requests of one stream in one call, with a gap of more than a segment between them
    skips = [100, seg - 50, 2 * seg + 999, 3 * seg + 100]
    counts = [seg - 200, 100, 40000, 4000]
    u, off = fe.rand([226] * 4, skips, counts)
    u = u.cpu().numpy()
    for k in range(4):
        assert np.array_equal(u[off[k]:off[k + 1]], ref[skips[k]:skips[k] + counts[k]]), k


def test_rand_deep_stream_positions(fe):
    """Segment indices above 255 need two jump polynomials (base-256 digits)."""
    seg = 4096 * 312
    for seed, skip, cnt in ((300, 17 * seg + 12345, 70000), (9, 257 * seg - 4000, 10000)):
        u, _ = fe.rand([seed], [skip], [cnt])
        assert np.array_equal(u.cpu().numpy(), _stream_slice(seed, skip, cnt)), (seed, skip)


# ---- a0 + a1: filtfilt -------------------------------------------------------------------------
def _filtfilt_cases():
    rng = np.random.default_rng(11)
    t = np.arange(70000) / 16000.0
    base = 0.3 * np.sin(2 * np.pi * 110 * t) + 0.05 * rng.standard_normal(t.shape[0]) + 0.02
    return [base[:48000], base[:32768], base[:20011], base[:19], base[:257] + 0.5, base[:65536],
            np.zeros(3000), 0.7 * np.ones(5000), base[:255], base[:256], base[:1000]]


@pytest.mark.parametrize("mode", ["scan", "sequential"])
def test_filtfilt_vs_scipy(fe, fe_seq, mode):
    f = fe if mode == "scan" else fe_seq
    xs = _filtfilt_cases()
    x, off = _ragged(xs, np.float64)
    y, fix = f.filtfilt(torch.from_numpy(x), off)
    y = y.cpu().numpy()
    for i, xi in enumerate(xs):
        ref = rp.highpass_filtfilt(rp.length_fixup(xi))
        got = y[fix[i]:fix[i + 1]]
        assert got.shape == ref.shape
        err = np.abs(got - ref).max()
        assert err <= 1e-6, "utt %d (L=%d): %g" % (i, len(xi), err)


def test_filtfilt_sequential_mode_is_tight(fe_seq):
    """One thread per utterance with scipy's operation order: agreement far below scipy's own
    round-off, which shows the scan's residual is conditioning, not a different filter."""
    xs = _filtfilt_cases()[:4]
    x, off = _ragged(xs, np.float64)
    y, fix = fe_seq.filtfilt(torch.from_numpy(x), off)
    y = y.cpu().numpy()
    for i, xi in enumerate(xs):
        ref = rp.highpass_filtfilt(rp.length_fixup(xi))
        assert np.abs(y[fix[i]:fix[i + 1]] - ref).max() <= 1e-12


def test_filtfilt_input_dtypes_agree(fe):
    pcm = synth_batch(make_manifest(1, 2, seed=5))
    x64, off = _ragged([pcm_to_float64(p) for p in pcm], np.float64)
    y64, _ = fe.filtfilt(torch.from_numpy(x64), off)
    y32, _ = fe.filtfilt(torch.from_numpy(x64.astype(np.float32)), off)
    y16, _ = fe.filtfilt(torch.cat(pcm), off)
    assert torch.equal(y64, y32) and torch.equal(y64, y16)     # int16/32768 is exact in f32 and f64


def test_filtfilt_too_short_raises(fe):
    with pytest.raises(ValueError):      # scipy: "length of the input vector x must be greater than padlen"
        fe.filtfilt(torch.zeros(10, dtype=torch.float64), [0, 10])


def test_filtfilt_long_form(fe):
    """60 s utterance (BASELINE config 4): 3751 chunks of carry."""
    rng = np.random.default_rng(2)
    x = 0.1 * rng.standard_normal(960000) + 0.2 * np.sin(np.arange(960000) * 2 * np.pi * 95 / 16000)
    y, fix = fe.filtfilt(torch.from_numpy(x), [0, 960000])
    ref = rp.highpass_filtfilt(rp.length_fixup(x))
    assert fix[-1] == 960001 and np.abs(y.cpu().numpy() - ref).max() <= 1e-6


# ---- a3 (+a4+a5): STFT and the fused mel kernel ------------------------------------------------
def _wav_cases():
    pcm = synth_batch(make_manifest(2, 2, seed=9))
    ws = [pcm_to_float64(p) * 0.96 for p in pcm]
    rng = np.random.default_rng(4)
    ws += [1e-6 * (rng.random(5000) - 0.5),            # pure dither (everything at the -100 dB floor)
           0.5 * np.sin(2 * np.pi * 1000 * np.arange(4097) / 16000.0),
           rng.standard_normal(1024) * 0.01, rng.standard_normal(700) * 0.1, rng.standard_normal(255)]
    return ws


def test_stft_mag_vs_pystft(fe):
    ws = _wav_cases()
    w, off = _ragged(ws, np.float32)
    mag, foff = fe.stft_mag(torch.from_numpy(w), off)
    mag = mag.cpu().numpy()
    for i, wi in enumerate(ws):
        ref = rp.pySTFT(wi.astype(np.float32).astype(np.float64)).T
        got = mag[foff[i]:foff[i + 1]]
        assert got.shape == ref.shape
        tol = 1e-5 * np.maximum(ref.max(axis=1, keepdims=True), 1e-30)
        assert np.all(np.abs(got - ref) <= tol), "utt %d: %g" % (i, (np.abs(got - ref) / tol).max())


def test_fused_mel_db_vs_oracle(fe):
    ws = _wav_cases()
    w, off = _ragged(ws, np.float32)
    mel, foff = fe.stft_mel_db(torch.from_numpy(w), off)
    mel = mel.cpu().numpy()
    worst = 0.0
    for i, wi in enumerate(ws):
        ref = rp.mel_db_normalize(rp.pySTFT(wi.astype(np.float32).astype(np.float64)).T)
        got = mel[foff[i]:foff[i + 1]]
        assert got.shape == ref.shape and got.dtype == np.float32
        worst = max(worst, np.abs(got - ref).max())
    assert worst <= 1e-4, worst
    # the pure-dither utterance sits on the clamp: S == (-100-16+100)/100
    sil = mel[foff[4]:foff[5]]
    assert np.allclose(sil, -0.16, atol=1e-6)


def test_fused_mel_frame_counts(fe):
    for L, T in ((48000 + 1, 188), (48128 + 1, 189), (32768 + 1, 129), (20011, 79), (1, 1), (255, 1), (256, 2)):
        mel, foff = fe.stft_mel_db(torch.zeros(L), [0, L])
        assert mel.shape == (T, 80) and foff[-1] == T


# ---- a7 + a8 + a9 ------------------------------------------------------------------------------
def _f0_cases():
    rng = np.random.default_rng(21)
    out = []
    for T in (1, 7, 8, 9, 10, 12, 127, 128, 129, 188, 300, 1000, 3751, 60000):   # 60000: past the warp kernel's planned tree
        f0 = rng.normal(5.0, 0.25, T).astype(np.float32)
        f0[rng.random(T) < 0.3] = np.float32(-1e10)
        out.append(f0)
    out.append(rng.normal(5.0, 0.25, 2048).astype(np.float32))  # all voiced, a power of two
    out.append(np.full(50, -1e10, np.float32))                 # all unvoiced -> nan stats, output untouched
    one = np.full(20, -1e10, np.float32)
    one[3] = 5.0                                               # single voiced frame -> std 0 -> nan
    out.append(one)
    return out


def test_f0_stats_and_normalization_bit_exact(fe):
    f0s = _f0_cases()
    f0, off = _ragged(f0s, np.float32)
    f0n, stats = fe.f0_normalize(torch.from_numpy(f0), off)
    f0n, stats = f0n.cpu().numpy(), stats.cpu().numpy()
    for i, fi in enumerate(f0s):
        idx, mean, std = rp.f0_stats(fi)
        with np.errstate(all="ignore"):
            ref = rp.speaker_normalization(fi, idx, mean, std).astype(np.float32)
        assert np.array_equal(stats[i], np.array([mean, std], np.float32), equal_nan=True), "stats of %d" % i
        assert np.array_equal(f0n[off[i]:off[i + 1]], ref, equal_nan=True), "utt %d" % i


def test_speaker_normalization_golden(fe, golden_dir):
    from speechsplit_b200 import utils
    k = np.load(os.path.join(golden_dir, "utils_kat.npz"))
    out = utils.speaker_normalization(k["sn_f0"], k["sn_f0"] != -1e10, k["sn_mean"], k["sn_std"])
    assert out.dtype == np.float64 and np.array_equal(out, k["sn_out"])


def test_quantize_numpy_golden(fe, golden_dir):
    from speechsplit_b200 import utils
    k = np.load(os.path.join(golden_dir, "utils_kat.npz"))
    x = k["q_in"]
    keep = x.copy()
    enc, idx = utils.quantize_f0_numpy(x)
    assert np.array_equal(x, keep)                                  # input not mutated (utils.py:49)
    assert enc.dtype == np.float32 and enc.shape == (x.shape[0], 257) and idx.dtype == np.int64
    assert np.array_equal(idx, k["q_idx"])
    assert np.array_equal(enc.argmax(1), k["q_enc_argmax"]) and np.array_equal(enc.sum(1), k["q_enc_sum"])
    ref_enc, ref_idx = rp.quantize_f0_numpy(x)
    assert np.array_equal(enc, ref_enc)
    with pytest.raises(AssertionError):
        utils.quantize_f0_numpy(np.array([0.2, 1.0001]))
    with pytest.raises(AssertionError):
        utils.quantize_f0_numpy(np.array([0.2, np.nan]))
    with pytest.raises(AssertionError):
        utils.quantize_f0_numpy(np.zeros((3, 3)))
    e0, i0 = utils.quantize_f0_numpy(np.zeros(0))
    assert e0.shape == (0, 257) and i0.shape == (0,)


def test_quantize_torch_golden(fe, golden_dir):
    from speechsplit_b200 import utils
    k = np.load(os.path.join(golden_dir, "utils_kat.npz"))
    x = torch.from_numpy(k["qt_in"]).cuda()
    enc, idx = utils.quantize_f0_torch(x)
    assert enc.is_cuda and enc.shape == (3, 192, 257) and enc.dtype == torch.float32 and idx.dtype == torch.int64
    assert np.array_equal(idx.cpu().numpy(), k["qt_idx"])
    assert np.array_equal(enc.argmax(-1).cpu().numpy(), k["qt_enc_argmax"])
    assert torch.all(enc.sum(-1) == 1)


def test_pystft_golden(fe, golden_dir):
    from speechsplit_b200 import utils
    k = np.load(os.path.join(golden_dir, "utils_kat.npz"))
    D = utils.pySTFT(k["stft_x"])
    assert D.dtype == np.float64 and D.shape == (513, 20)
    ref = k["stft_D"].astype(np.float64)
    assert np.all(np.abs(D - ref) <= 3e-6 * ref.max(axis=0, keepdims=True))


def test_collate_matches_data_loader(fe):
    """data_loader.py:101-128: crop, clip [0,1], zero-pad to 192, F0 pad -1e10; then solver.py:162."""
    rng = np.random.default_rng(8)
    Ts = [135, 200, 70]
    mel = (rng.random((sum(Ts), 80)) * 1.3 - 0.15).astype(np.float32)
    f0 = rng.random(sum(Ts)).astype(np.float32)
    f0[rng.random(sum(Ts)) < 0.3] = -1e10
    foff = np.concatenate([[0], np.cumsum(Ts)])
    utt, left, ln = [0, 1, 2, 1], [3, 50, 0, 0], [128, 64, 70, 100]
    melsp, pitch, onehot, bins = fe.collate(mel, f0, foff, utt, left, ln, 192)
    for i in range(4):
        a = mel[foff[utt[i]] + left[i]: foff[utt[i]] + left[i] + ln[i]]
        c = f0[foff[utt[i]] + left[i]: foff[utt[i]] + left[i] + ln[i]]
        a_pad = np.pad(np.clip(a, 0, 1), ((0, 192 - ln[i]), (0, 0)), "constant")
        c_pad = np.pad(c[:, None], ((0, 192 - ln[i]), (0, 0)), "constant", constant_values=-1e10)
        assert np.array_equal(melsp[i].cpu().numpy(), a_pad)
        assert np.array_equal(pitch[i].cpu().numpy(), c_pad.astype(np.float32))
        enc, idx = rp.quantize_f0_numpy(c_pad[:, 0])
        assert np.array_equal(bins[i].cpu().numpy(), idx) and np.array_equal(onehot[i].cpu().numpy(), enc)
    with pytest.raises(Exception):
        fe.collate(mel, f0, foff, [0], [100], [128], 192)        # crop runs past the utterance


def test_loader_mirror_on_gpu(fe, tmp_path):
    """speechsplit_b200.data_loader (get_loader -> Utterances / MyCollator, features resident in HBM, one
    ssfe_collate launch per batch) against the reference's collator loop (data_loader.py:101-128) on the
    same numpy seed: identical melsp / spk_emb / pitch / len_org, as CUDA tensors of the shapes and dtypes
    solver.py:142 unpacks; the optional one-hot equals quantize_f0_torch of the padded pitch (solver.py:162)."""
    from types import SimpleNamespace

    from speechsplit_b200.data_loader import get_loader, make_metadata
    from test_host_checks import _feature_tree, _reference_collate

    _feature_tree(tmp_path, {"p225": [150, 400], "p226": [200], "p227": [135, 140], "p228": [129]})
    make_metadata(str(tmp_path / "spmel"), verbose=False)
    hp = SimpleNamespace(root_dir=str(tmp_path / "spmel"), feat_dir=str(tmp_path / "raptf0"), mode="train",
                         batch_size=16, shuffle=True, num_workers=0, samplier=8, min_len_seq=64, max_len_seq=128,
                         max_len_pad=192)
    loader = get_loader(hp, frontend=fe, want_onehot=True)
    ds = loader.dataset
    assert len(loader) == (4 * 8) // 16
    torch.manual_seed(5)
    np.random.seed(21)
    n_launch = fe.launch_count
    got, onehots = [], []
    for batch in loader:
        got.append(batch)
        onehots.append(loader.collate_fn.last_onehot)
    assert fe.launch_count - n_launch <= 5 * len(got)          # the collate kernel + its metadata copies
    order = list(loader.sampler.sample_idx_array.numpy())       # the shuffled pass the loader just made
    assert sorted(order) == sorted(list(range(4)) * 8) and order != sorted(order)
    np.random.seed(21)
    for b, (batch, (onehot, bins)) in enumerate(zip(got, onehots)):
        want = _reference_collate([tuple(ds[i]) for i in order[16 * b: 16 * b + 16]], 64, 128, 192)
        for g, w, shape, dt in zip(batch, want, [(16, 192, 80), (16, 82), (16, 192, 1), (16,)],
                                   [torch.float32, torch.float32, torch.float32, torch.int64]):
            assert g.is_cuda and tuple(g.shape) == shape and g.dtype == dt
            assert torch.equal(g.cpu(), w)
        enc, idx = rp.quantize_f0_numpy(want[2].numpy().reshape(-1))
        assert np.array_equal(onehot.cpu().numpy().reshape(-1, 257), enc)
        assert np.array_equal(bins.cpu().numpy().reshape(-1), idx)


def test_collator_golden_on_gpu(fe, tmp_path, golden_dir):
    """MyCollator on the GPU == the batch the reference's OWN data_loader.MyCollator produced for the same
    items and numpy seed (tests/golden/collate.npz): bit-exact melsp / spk_emb / pitch / len_org."""
    from types import SimpleNamespace

    from speechsplit_b200.data_loader import MyCollator, Utterances, make_metadata
    from test_host_checks import _golden_tree

    g = _golden_tree(tmp_path, golden_dir)
    make_metadata(str(tmp_path / "spmel"), verbose=False)
    ds = Utterances(str(tmp_path / "spmel"), str(tmp_path / "raptf0"), "train", frontend=fe)
    col = MyCollator(SimpleNamespace(min_len_seq=64, max_len_seq=128, max_len_pad=192), ds)
    np.random.seed(int(g["seed"]))
    batch = col([ds[int(i)] for i in g["order"]])
    for got, name in zip(batch, ("melsp", "spk_emb", "pitch", "len_org")):
        assert got.is_cuda and got.cpu().numpy().dtype == g[name].dtype
        assert np.array_equal(got.cpu().numpy(), g[name]), name


def test_demo_notebook_inputs(fe, golden_dir):
    """demo.ipynb:36-50 on the reference's own assets/demo.pkl (real VCTK features made by the original
    pipeline): utils.pad_seq_to_2 + utils.quantize_f0_numpy give what the reference's utils give."""
    from speechsplit_b200 import utils
    g = np.load(os.path.join(golden_dir, "demo_kat.npz"))
    for k in range(int(g["n"])):
        f0, mel = g["f0%d" % k], g["mel%d" % k]
        mel_pad, len_pad = utils.pad_seq_to_2(mel[np.newaxis, :, :], 192)
        f0_pad = np.pad(f0, (0, 192 - len(f0)), "constant", constant_values=(0, 0))
        enc, idx = utils.quantize_f0_numpy(f0_pad)
        assert np.array_equal(mel_pad, g["mel_pad%d" % k]) and len_pad == int(g["len_pad%d" % k])
        assert enc.dtype == np.float32 and enc.shape == (192, 257) and idx.dtype == np.int64
        assert np.array_equal(idx, g["idx%d" % k])
        assert np.array_equal(enc.argmax(1), g["enc_argmax%d" % k]) and np.array_equal(enc.sum(1), g["enc_sum%d" % k])
