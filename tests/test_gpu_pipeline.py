"""GPU parity of the whole hot loop (ssfe_extract / ssfe_extract_host through the C ABI) against
the reference-generated golden vectors and the oracle pipeline.

Tolerances (BASELINE.json north_star): normalised mel <= 1e-4 abs; quantised-F0 bins and voicing
flags identical on >= 99.9 % of frames; F0 within 1 cent on frames voiced in both; frame counts
identical.  With the sequential filtfilt validation mode the RAPT input is bit-identical to the
oracle's, and every F0 / bin must match exactly."""
import os

import numpy as np
import pytest
import torch
from numpy.random import RandomState

from oracle import ref_pipeline as rp
from speechsplit_b200.corpus import make_manifest, pcm_to_float64, synth_batch

pytestmark = pytest.mark.gpu
UNV = np.float32(-1e10)


def _golden_batch(golden_dir, names):
    pcm, meta = [], []
    for name in names:
        g = np.load(os.path.join(golden_dir, name))
        spk, gender = str(g["spk"]), str(g["gender"])
        skip = 0
        for k in range(int(g["n"])):
            p = g["pcm%d" % k]
            pcm.append(p)
            meta.append(dict(spk=spk, gender=gender, skip=skip, S=g["S%d" % k], f0=g["f0_rapt%d" % k],
                             f0n=g["f0_norm%d" % k], bins=g["bins%d" % k].astype(np.int64),
                             wav=g["wav0"] if k == 0 else None))
            L = p.shape[0]
            skip += L + (1 if L % 256 == 0 else 0)
    return pcm, meta


def _extract(fe, pcm, meta, want, dtype=torch.int16):
    off = np.concatenate([[0], np.cumsum([len(p) for p in pcm])]).astype(np.int64)
    x = torch.from_numpy(np.concatenate(pcm))
    if dtype != torch.int16:
        x = (x.double() / 32768.0).to(dtype)
    lo = [50.0 if m["gender"] == "M" else 100.0 for m in meta]
    hi = [250.0 if m["gender"] == "M" else 600.0 for m in meta]
    seed = [int(m["spk"][1:]) for m in meta]
    skip = [m["skip"] for m in meta]
    return fe.extract(x, off, lo, hi, seed, skip, want=want)


def _extract_out(fe, pcm, meta, want, out):
    off = np.concatenate([[0], np.cumsum([len(p) for p in pcm])]).astype(np.int64)
    x = torch.from_numpy(np.concatenate(pcm))
    lo = [50.0 if m["gender"] == "M" else 100.0 for m in meta]
    hi = [250.0 if m["gender"] == "M" else 600.0 for m in meta]
    return fe.extract(x, off, lo, hi, [int(m["spk"][1:]) for m in meta], [m["skip"] for m in meta], want=want, out=out)


def _check_against(res, meta, exact_f0):
    fo, fx = res["frame_offsets"], res["fixed_offsets"]
    mel = res["mel"].cpu().numpy()
    f0n = res["f0_norm"].cpu().numpy()
    f0 = res["f0_raw"].cpu().numpy()
    bins = res["bins"].cpu().numpy()
    onehot = res["onehot"].cpu().numpy()
    n_fr = n_same = 0
    worst_mel = worst_cent = 0.0
    for i, m in enumerate(meta):
        S = mel[fo[i]:fo[i + 1]]
        assert S.shape == m["S"].shape and S.dtype == np.float32            # len(S) == len(f0_rapt), :69
        worst_mel = max(worst_mel, float(np.abs(S - m["S"]).max()))
        g_f0, g_bins, g_f0n = f0[fo[i]:fo[i + 1]], bins[fo[i]:fo[i + 1]], f0n[fo[i]:fo[i + 1]]
        assert g_f0.shape == m["f0"].shape
        same = (g_bins == m["bins"]) & ((g_f0 == UNV) == (m["f0"] == UNV))
        n_fr += same.size
        n_same += int(same.sum())
        both = (g_f0 != UNV) & (m["f0"] != UNV)
        if both.any():
            worst_cent = max(worst_cent, float(1731.234 * np.abs(g_f0[both] - m["f0"][both]).max()))
        if exact_f0:
            assert np.array_equal(g_f0, m["f0"]) and np.array_equal(g_f0n, m["f0n"], equal_nan=True)
            assert np.array_equal(g_bins, m["bins"])
        # one-hot consistent with bins: exactly quantize_f0_numpy's encoding
        enc = onehot[fo[i]:fo[i + 1]]
        assert enc.shape == (len(g_bins), 257) and np.array_equal(enc.argmax(1), g_bins) and np.all(enc.sum(1) == 1)
        if m["wav"] is not None and "wav64" in res:
            w = res["wav64"].cpu().numpy()[fx[i]:fx[i + 1]]
            assert np.abs(w - m["wav"]).max() <= (1e-12 if exact_f0 else 1e-6)
    assert worst_mel <= 1e-4, worst_mel
    assert n_same >= 0.999 * n_fr, (n_same, n_fr)
    assert worst_cent <= 1.0, worst_cent
    return worst_mel, n_same / n_fr, worst_cent


WANT = ("mel", "f0_norm", "f0_raw", "bins", "onehot", "wav", "wav64")
NAMES = ["pipeline_p226.npz", "pipeline_p225.npz"]


def test_extract_golden_scan_mode(fe, golden_dir):
    pcm, meta = _golden_batch(golden_dir, NAMES)
    res = _extract(fe, pcm, meta, WANT)
    _check_against(res, meta, exact_f0=False)


def test_extract_golden_production_path(fe, golden_dir):
    """Without the fp64 wav among the outputs ssfe_extract takes its production path: the dither term travels
    as float (mt_walk_kernel<true>) and the backward filter pass writes the padded f32 segments directly."""
    pcm, meta = _golden_batch(golden_dir, NAMES)
    res = _extract(fe, pcm, meta, ("mel", "f0_norm", "f0_raw", "bins", "onehot"))
    _check_against(res, meta, exact_f0=False)
    ref = _extract(fe, pcm, meta, WANT)
    # float dither vs raw word pairs: the f32 wav may differ in the last bit on a few samples per million, far
    # below what moves a mel value
    assert (res["mel"] - ref["mel"]).abs().max().item() <= 2e-6
    assert (res["bins"] == ref["bins"]).float().mean().item() >= 0.999


def test_extract_long_form_60s_whole_loop(fe):
    """BASELINE configs[3] shape: the whole loop on 60.000 s utterances (L = 960 000 -> 960 001 by the :52-53
    append, 3 751 frames; the filter carry takes the warp-scan path, the Viterbi pass 3 751 steps), one male and
    one female speaker, against the oracle - same gates as everywhere."""
    metas = make_manifest(2, 1, seed=4, fixed_len=960000)
    pcm = [p.numpy() for p in synth_batch(metas)]
    meta = []
    for m, p in zip(metas, pcm):
        S, f0n, st = rp.extract_utterance(pcm_to_float64(p), m.gender, RandomState(m.spk_id), want_stages=True)
        assert S.shape == (3751, 80)
        meta.append(dict(spk=m.spk, gender=m.gender, skip=0, S=S, f0=st["f0_rapt"], f0n=f0n,
                         bins=rp.quantize_f0_numpy(f0n)[1], wav=None))
    res = _extract(fe, pcm, meta, ("mel", "f0_norm", "f0_raw", "bins", "onehot"))
    assert res["frame_offsets"][-1] == 2 * 3751 and res["fixed_offsets"][-1] == 2 * 960001
    worst_mel, frac, cents = _check_against(res, meta, exact_f0=False)
    print("60 s whole loop: mel %.2e  identical bins %.5f  worst F0 %.3f cent" % (worst_mel, frac, cents))


def test_extract_golden_sequential_mode_is_exact(fe_seq, golden_dir):
    pcm, meta = _golden_batch(golden_dir, NAMES)
    res = _extract(fe_seq, pcm, meta, WANT)
    _check_against(res, meta, exact_f0=True)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_extract_input_dtypes(fe, golden_dir, dtype):
    pcm, meta = _golden_batch(golden_dir, NAMES[:1])
    a = _extract(fe, pcm, meta, ("mel", "f0_norm", "bins"))
    b = _extract(fe, pcm, meta, ("mel", "f0_norm", "bins"), dtype=dtype)
    for k in ("mel", "f0_norm", "bins"):
        assert torch.equal(a[k], b[k]), k


def test_extract_batch_order_and_split_invariance(fe, golden_dir):
    """Utterances are independent given (seed, skip): shuffling the batch or running each utterance
    alone gives bit-identical results - the property the multi-GPU sharding relies on."""
    pcm, meta = _golden_batch(golden_dir, NAMES)
    full = _extract(fe, pcm, meta, ("mel", "f0_norm", "bins"))
    fo = full["frame_offsets"]
    perm = [3, 0, 4, 2, 1]
    sh = _extract(fe, [pcm[i] for i in perm], [meta[i] for i in perm], ("mel", "f0_norm", "bins"))
    so = sh["frame_offsets"]
    for k, i in enumerate(perm):
        for name in ("mel", "f0_norm", "bins"):
            assert torch.equal(sh[name][so[k]:so[k + 1]], full[name][fo[i]:fo[i + 1]]), (name, i)
    for i in range(len(pcm)):
        one = _extract(fe, [pcm[i]], [meta[i]], ("mel", "f0_norm", "bins"))
        for name in ("mel", "f0_norm", "bins"):
            assert torch.equal(one[name], full[name][fo[i]:fo[i + 1]]), (name, i)


def test_extract_onehot_zero_stream_paths(fe, golden_dir, monkeypatch):
    """The one-hot rows are written as a zero stream on a side stream beside the RAPT kernels plus one 1.0 per
    row from the normalisation kernel (ssfe_extract, api.cu).  That path, the single-kernel path it replaced
    (SSFE_ONEHOT_EARLY=0) and the fallback for a buffer that is not 16-byte aligned must give the same bits,
    whatever the buffer held before, for row counts of every residue mod 4."""
    from speechsplit_b200 import FrontEnd
    pcm, meta = _golden_batch(golden_dir, NAMES)
    monkeypatch.setenv("SSFE_ONEHOT_EARLY", "0")
    late = FrontEnd(0)
    monkeypatch.delenv("SSFE_ONEHOT_EARLY")
    try:
        seen = set()
        for k in range(1, len(pcm) + 1):
            ref = _extract(late, pcm[:k], meta[:k], ("mel", "f0_norm", "bins", "onehot"))
            T = int(ref["frame_offsets"][-1])
            seen.add(T % 4)
            assert torch.equal(ref["onehot"].argmax(1), ref["bins"]) and bool((ref["onehot"].sum(1) == 1).all())
            for shift in (0, 1):            # shift 1: rows start 4 bytes off a 16-byte boundary
                buf = torch.full((T * 257 + 4,), 7.0, dtype=torch.float32, device=ref["onehot"].device)
                oh = buf[shift:shift + T * 257].view(T, 257)
                for _ in range(2):          # twice: the second call finds the first call's ones in the buffer
                    got = _extract_out(fe, pcm[:k], meta[:k], ("mel", "f0_norm", "bins", "onehot"), dict(onehot=oh))
                    assert got["onehot"].data_ptr() == oh.data_ptr()
                    for name in ("mel", "f0_norm", "bins", "onehot"):
                        assert torch.equal(got[name], ref[name]), (name, k, shift)
                assert bool((buf[:shift] == 7.0).all()) and bool((buf[shift + T * 257:] == 7.0).all())
        assert len(seen) >= 2
    finally:
        late.close()


def test_extract_host_matches_device_path(fe, golden_dir):
    pcm, meta = _golden_batch(golden_dir, NAMES)
    res = _extract(fe, pcm, meta, ("mel", "f0_norm", "bins"))
    off = np.concatenate([[0], np.cumsum([len(p) for p in pcm])]).astype(np.int64)
    lo = [50.0 if m["gender"] == "M" else 100.0 for m in meta]
    hi = [250.0 if m["gender"] == "M" else 600.0 for m in meta]
    h = fe.extract_host(np.concatenate(pcm), off, lo, hi, [int(m["spk"][1:]) for m in meta],
                        [m["skip"] for m in meta])
    assert np.array_equal(h["mel"], res["mel"].cpu().numpy())
    assert np.array_equal(h["f0_norm"], res["f0_norm"].cpu().numpy(), equal_nan=True)
    assert np.array_equal(h["bins"], res["bins"].cpu().numpy())


def test_extract_host_many_sub_batches(golden_dir, monkeypatch):
    """The pipelined host path (up-front dither, double-buffered H2D / kernels / D2H) with the
    sub-batch size forced down so that every utterance is its own sub-batch."""
    from speechsplit_b200 import FrontEnd
    monkeypatch.setenv("SSFE_HOST_CHUNK_SAMPLES", "1000")
    f2 = FrontEnd(0)
    try:
        pcm, meta = _golden_batch(golden_dir, NAMES)
        ref = _extract(f2, pcm, meta, ("mel", "f0_norm", "bins"))
        off = np.concatenate([[0], np.cumsum([len(p) for p in pcm])]).astype(np.int64)
        lo = [50.0 if m["gender"] == "M" else 100.0 for m in meta]
        hi = [250.0 if m["gender"] == "M" else 600.0 for m in meta]
        for _ in range(2):      # twice: slot reuse and the dither-buffer hand-over between calls
            h = f2.extract_host(np.concatenate(pcm), off, lo, hi, [int(m["spk"][1:]) for m in meta],
                                [m["skip"] for m in meta])
            assert np.array_equal(h["mel"], ref["mel"].cpu().numpy())
            assert np.array_equal(h["f0_norm"], ref["f0_norm"].cpu().numpy(), equal_nan=True)
            assert np.array_equal(h["bins"], ref["bins"].cpu().numpy())
    finally:
        f2.close()


def test_extract_vs_oracle_small_corpus(fe):
    """A fresh synthetic mini-corpus (6 speakers x 4 files) against the oracle pipeline."""
    metas = make_manifest(6, 4, seed=17)
    pcm = [p.numpy() for p in synth_batch(metas)]
    meta, cur, prng, skip = [], None, None, 0
    for m, p in zip(metas, pcm):
        if cur != m.spk:
            cur, prng, skip = m.spk, RandomState(m.spk_id), 0
        S, f0n, st = rp.extract_utterance(pcm_to_float64(p), m.gender, prng, want_stages=True)
        meta.append(dict(spk=m.spk, gender=m.gender, skip=skip, S=S, f0=st["f0_rapt"], f0n=f0n,
                         bins=rp.quantize_f0_numpy(f0n)[1], wav=None))
        skip += len(p) + (1 if len(p) % 256 == 0 else 0)
    res = _extract(fe, pcm, meta, WANT)
    worst_mel, frac, cents = _check_against(res, meta, exact_f0=False)
    print("mini-corpus: mel %.2e, identical frames %.5f, worst %.3f cent" % (worst_mel, frac, cents))


def test_extract_errors(fe):
    x = torch.zeros(500, dtype=torch.int16)
    with pytest.raises(ValueError):
        fe.extract(x, [0, 500], [50.0], [250.0], [226], [0])             # too short for get_f0
    with pytest.raises(ValueError):
        fe.extract(torch.zeros(5000, dtype=torch.int16), [0, 5000], [70.0], [250.0], [226], [0])   # not M / F


def test_extract_refuses_bad_buffers(fe):
    """Offsets that leave the sample buffer and output buffers of the wrong shape never reach the C ABI."""
    x = torch.zeros(5000, dtype=torch.int16)
    with pytest.raises(ValueError):
        fe.extract(x, [0, 5001], [50.0], [250.0], [226], [0])
    with pytest.raises(ValueError):
        fe.extract_host(x.numpy(), [0, 5001], [50.0], [250.0], [226], [0])
    with pytest.raises(TypeError):
        fe.extract(torch.zeros(5000, dtype=torch.int32), [0, 5000], [50.0], [250.0], [226], [0])
    with pytest.raises(ValueError):
        fe.extract(x, [0, 5000], [50.0], [250.0], [226], [0],
                   out=dict(mel=torch.empty((19, 80), device="cuda")))       # 5000 samples are 20 frames
    with pytest.raises(ValueError):
        fe.extract_host(x.numpy(), [0, 5000], [50.0], [250.0], [226], [0],
                        out=dict(mel=np.empty((20, 80), np.float64), f0_norm=np.empty(20, np.float32)))


def test_extract_host_error_in_later_sub_batch(golden_dir, monkeypatch):
    """A sub-batch that fails after earlier ones were queued: the error surfaces as the reference's
    ValueError, nothing is left in flight on the caller's buffers, and the context stays usable."""
    from speechsplit_b200 import FrontEnd
    monkeypatch.setenv("SSFE_HOST_CHUNK_SAMPLES", "1000")
    f2 = FrontEnd(0)
    try:
        pcm, meta = _golden_batch(golden_dir, NAMES)
        lo = [50.0 if m["gender"] == "M" else 100.0 for m in meta]
        hi = [250.0 if m["gender"] == "M" else 600.0 for m in meta]
        seed, skip = [int(m["spk"][1:]) for m in meta], [m["skip"] for m in meta]
        off = np.concatenate([[0], np.cumsum([len(p) for p in pcm])]).astype(np.int64)
        good = f2.extract_host(np.concatenate(pcm), off, lo, hi, seed, skip)
        # same batch with a 500-sample utterance of another speaker appended: too short for get_f0
        bad_pcm = pcm + [np.zeros(500, np.int16)]
        bad_off = np.concatenate([off, [off[-1] + 500]]).astype(np.int64)
        with pytest.raises(ValueError):
            f2.extract_host(np.concatenate(bad_pcm), bad_off, lo + [50.0], hi + [250.0], seed + [999], skip + [0])
        again = f2.extract_host(np.concatenate(pcm), off, lo, hi, seed, skip)
        for k in ("mel", "f0_norm", "bins"):
            assert np.array_equal(again[k], good[k], equal_nan=True), k
    finally:
        f2.close()


def test_make_spect_f0_script_drop_in(tmp_path):
    """The script form (speechsplit_b200.make_spect_f0) against the reference's loop run on the oracle:
    a tree of 16-bit mono WAVs + spk2gen.pkl in, spmel/ and raptf0/ trees of NPY v1.0 '<f4' out
    (make_spect_f0.py:19-31,47-74), files visited in sorted() order so that each speaker's dither
    stream continues across its files."""
    import pickle
    import wave

    from speechsplit_b200.make_spect_f0 import make_spect_f0

    metas = make_manifest(2, 3, seed=23)
    pcm = [p.numpy() for p in synth_batch(metas)]
    root, out, out_f0 = tmp_path / "wavs", tmp_path / "spmel", tmp_path / "raptf0"
    spk2gen = {}
    for k, (m, p) in enumerate(zip(metas, pcm)):
        (root / m.spk).mkdir(parents=True, exist_ok=True)
        spk2gen[m.spk] = m.gender
        # names chosen so that sorted() order differs from creation order within a speaker
        with wave.open(str(root / m.spk / ("%s_%03d.wav" % (m.spk, 900 - k))), "wb") as w:
            w.setnchannels(1), w.setsampwidth(2), w.setframerate(16000)
            w.writeframes(p.astype("<i2").tobytes())
    with open(tmp_path / "spk2gen.pkl", "wb") as f:
        pickle.dump(spk2gen, f)

    make_spect_f0(str(root), str(out), str(out_f0), str(tmp_path / "spk2gen.pkl"), verbose=False)

    worst, same, total = 0.0, 0, 0
    for spk in sorted(spk2gen):
        files = sorted(os.listdir(root / spk))
        prng = RandomState(int(spk[1:]))
        for fname in files:
            with wave.open(str(root / spk / fname), "rb") as w:
                x = np.frombuffer(w.readframes(w.getnframes()), dtype="<i2")
            S_ref, f0n_ref = rp.extract_utterance(pcm_to_float64(x), spk2gen[spk], prng)
            mel_path, f0_path = out / spk / (fname[:-4] + ".npy"), out_f0 / spk / (fname[:-4] + ".npy")
            with open(mel_path, "rb") as f:
                head = f.read(10)
            assert head[:6] == b"\x93NUMPY" and head[6:8] == b"\x01\x00"      # NPY v1.0 (allow_pickle=False)
            S, f0n = np.load(mel_path, allow_pickle=False), np.load(f0_path, allow_pickle=False)
            assert S.dtype == np.float32 and f0n.dtype == np.float32
            assert S.shape == S_ref.shape and f0n.shape == f0n_ref.shape and S.shape[0] == f0n.shape[0]   # :69
            worst = max(worst, float(np.abs(S - S_ref).max()))
            b, b_ref = rp.quantize_f0_numpy(f0n)[1], rp.quantize_f0_numpy(f0n_ref.astype(np.float32))[1]
            same += int((b == b_ref).sum())
            total += b.size
    assert worst <= 1e-4, worst
    assert same / total >= 0.999, (same, total)


def test_extract_large_corpus_properties(fe):
    """One GPU's share of the benchmark corpus at N=8 (14 speakers x 400 utterances, 16 800 audio-s,
    ~1.05 M frames) through size-independent properties: results do not depend on batch order or on how
    the corpus is split over calls (bitwise), the device and the host entry points agree, frame counts
    follow make_spect_f0.py:69, every one-hot row has exactly one 1 at its bin, and a sample of
    utterances matches the oracle within the north_star tolerances."""
    from speechsplit_b200.sharding import dither_skips
    metas = make_manifest(14, 400, seed=0)
    pcm = synth_batch(metas, device="cuda")
    lens = np.array([m.length for m in metas])
    skips = dither_skips([m.spk for m in metas], lens)
    lo = np.array([50.0 if m.gender == "M" else 100.0 for m in metas], np.float32)
    hi = np.array([250.0 if m.gender == "M" else 600.0 for m in metas], np.float32)
    seed = np.array([m.spk_id for m in metas], np.uint32)

    def run(idx, want=("mel", "f0_norm", "bins")):
        off = np.concatenate([[0], np.cumsum(lens[idx])]).astype(np.int64)
        x = torch.cat([pcm[i] for i in idx])
        return fe.extract(x, off, lo[idx], hi[idx], seed[idx], skips[idx], want=want)

    n = len(metas)
    whole = run(np.arange(n), want=("mel", "f0_norm", "bins", "onehot"))
    fo = whole["frame_offsets"]
    fixed = lens + (lens % 256 == 0)                                  # make_spect_f0.py:52-53
    assert np.array_equal(np.diff(fo), -(-fixed // 256))                # T = ceil(L' / 256) = len(f0_rapt), :69
    assert (lens % 256 == 0).sum() >= 80                               # the append path is exercised
    bins = whole["bins"]
    assert int(bins.min()) >= 0 and int(bins.max()) <= 256
    onehot = whole["onehot"]
    assert torch.equal(onehot.argmax(1), bins) and bool((onehot.sum(1) == 1).all())
    voiced = float((bins > 0).float().mean())
    assert 0.2 < voiced < 0.95, voiced

    # a permutation, processed as three separate calls, must reproduce every utterance bit for bit
    rng = np.random.default_rng(1)
    perm = rng.permutation(n)
    pos = 0
    for part in np.array_split(perm, 3):
        res = run(part)
        pfo = res["frame_offsets"]
        for k in rng.choice(len(part), 40, replace=False):
            i = part[k]
            assert torch.equal(res["mel"][pfo[k]:pfo[k + 1]], whole["mel"][fo[i]:fo[i + 1]]), i
            assert torch.equal(res["bins"][pfo[k]:pfo[k + 1]], whole["bins"][fo[i]:fo[i + 1]]), i
        pos += len(part)

    # a few utterances alone: a batch that cannot fill the GPU takes filter tiles of 8 chunks instead of 32 and the
    # single-round paths of the carry scan and the dot-product local passes - the bits must not notice
    for i in (0, 1234, 5599):
        one = run(np.array([i]))
        assert torch.equal(one["mel"], whole["mel"][fo[i]:fo[i + 1]]), i
        assert torch.equal(one["f0_norm"], whole["f0_norm"][fo[i]:fo[i + 1]]), i
        assert torch.equal(one["bins"], whole["bins"][fo[i]:fo[i + 1]]), i

    # host entry point == device entry point
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    host = fe.extract_host(torch.cat(pcm).cpu(), off, lo, hi, seed, skips)
    assert np.array_equal(host["mel"], whole["mel"].cpu().numpy())
    assert np.array_equal(host["bins"], whole["bins"].cpu().numpy())

    # a sample against the oracle (each with its speaker's stream advanced to the utterance)
    worst, same, total = 0.0, 0, 0
    for i in (0, 399, 400, 2801, 5599):
        m = metas[i]
        prng = RandomState(m.spk_id)
        left = int(skips[i])
        while left:
            step = min(left, 10_000_000)
            prng.rand(step)
            left -= step
        S, f0n = rp.extract_utterance(pcm_to_float64(pcm[i].cpu().numpy()), m.gender, prng)
        worst = max(worst, float(np.abs(whole["mel"][fo[i]:fo[i + 1]].cpu().numpy() - S).max()))
        b_ref = rp.quantize_f0_numpy(f0n.astype(np.float32))[1]
        b = whole["bins"][fo[i]:fo[i + 1]].cpu().numpy()
        same += int((b == b_ref).sum())
        total += b.size
    assert worst <= 1e-4, worst
    assert same >= 0.999 * total, (same, total)
