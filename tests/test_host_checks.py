"""Host-side argument checks of speechsplit_b200.frontend (no GPU needed).

The C ABI takes plain pointers and sizes (include/ssfe.h), so everything the reference gets for free
from numpy indexing - offsets that leave the buffer, output buffers of the wrong shape or dtype - has to
be refused before a pointer crosses the boundary."""
import os

import numpy as np
import pytest
import torch

from speechsplit_b200 import _lib as L
from speechsplit_b200.frontend import _DT, _NP_DT, _check_out, _check_ragged, _wav_dtype


def test_ragged_offsets_accept_valid():
    _check_ragged(np.array([0, 48000, 96001], np.int64), 96001)
    _check_ragged(np.array([0], np.int64), 0)                      # empty batch
    _check_ragged(np.array([0, 0, 10], np.int64), 10)              # empty utterance: the C side decides


@pytest.mark.parametrize("off, numel", [
    (np.array([0, 48000, 96002], np.int64), 96001),                # runs past the buffer
    (np.array([0, 500, 400], np.int64), 1000),                     # decreasing
    (np.array([-1, 400], np.int64), 1000),                         # negative start
    (np.zeros((2, 2), np.int64), 1000),                            # not 1-D
    (np.zeros(0, np.int64), 1000),                                 # no offsets at all
])
def test_ragged_offsets_refuse_invalid(off, numel):
    with pytest.raises(ValueError):
        _check_ragged(off, numel)


def test_wav_dtype_table():
    assert _wav_dtype(_DT, torch.int16) == L.I16 and _wav_dtype(_DT, torch.float64) == L.F64
    assert _wav_dtype(_NP_DT, np.dtype(np.float32)) == L.F32
    with pytest.raises(TypeError):
        _wav_dtype(_DT, torch.float16)
    with pytest.raises(TypeError):
        _wav_dtype(_NP_DT, np.dtype(np.int32))


def test_output_buffer_checks_host():
    _check_out("mel", np.empty((7, 80), np.float32), (7, 80), torch.float32)
    _check_out("bins", torch.empty(7, dtype=torch.int64), (7,), torch.int64)
    bad = [
        ("mel", np.empty((7, 80), np.float64), (7, 80), torch.float32),          # dtype
        ("mel", np.empty((8, 80), np.float32), (7, 80), torch.float32),          # shape
        ("mel", np.empty((80, 7), np.float32).T, (7, 80), torch.float32),        # not C-contiguous
        ("f0_norm", torch.empty(14)[::2], (7,), torch.float32),                  # strided tensor
        ("bins", [0] * 7, (7,), torch.int64),                                    # not an array
    ]
    for name, a, shape, dt in bad:
        with pytest.raises(ValueError):
            _check_out(name, a, shape, dt)
    ro = np.empty(7, np.float32)
    ro.flags.writeable = False
    with pytest.raises(ValueError):
        _check_out("f0_norm", ro, (7,), torch.float32)


def test_output_buffer_checks_device_side_refuses_cpu():
    # a device output must live on the context's GPU: a CPU tensor (or numpy array) is refused
    dev = torch.device("cuda", 0)
    with pytest.raises(ValueError):
        _check_out("mel", torch.empty((7, 80)), (7, 80), torch.float32, dev)
    with pytest.raises(ValueError):
        _check_out("mel", np.empty((7, 80), np.float32), (7, 80), torch.float32, dev)


class _RecordingFrontEnd:
    """Stands in for FrontEnd.extract_host: records every call and returns frame-shaped placeholders, so
    that the host logic of the script form (file order, stream positions, batching, NPY trees) is checked
    without a GPU.  It computes nothing - the arithmetic is covered by the -m gpu tests."""

    def __init__(self):
        self.calls = []

    def extract_host(self, x, off, lo, hi, seed, skip, want_bins=True):
        off = np.asarray(off)
        frames = (np.diff(off) + (np.diff(off) % 256 == 0) + 256) // 256
        fo = np.concatenate([[0], np.cumsum(frames)])
        self.calls.append(dict(n=len(lo), dtype=x.dtype, lo=list(lo), hi=list(hi), seed=list(seed), skip=list(skip)))
        tag = np.repeat(np.arange(len(lo), dtype=np.float32), frames)
        return dict(mel=np.repeat(tag[:, None], 80, 1), f0_norm=tag.copy(), frame_offsets=fo)


def test_script_form_host_logic(tmp_path, monkeypatch):
    """make_spect_f0.py:19-31,47-55,69-74: sorted speaker / file order, seed int(spk[1:]), the speaker's
    stream position advancing by the post-append length, one NPY pair per WAV - and ValueError for a
    gender that is neither 'M' nor 'F' (:45)."""
    import pickle
    import wave

    from speechsplit_b200 import make_spect_f0 as script

    lens = {"p300": [5120, 4000, 7000], "p225": [6000, 2560]}
    root = tmp_path / "wavs"
    for spk, ls in lens.items():
        (root / spk).mkdir(parents=True)
        for k, n in enumerate(ls):
            with wave.open(str(root / spk / ("%s_%03d.wav" % (spk, 9 - k))), "wb") as w:   # sorted() reverses k
                w.setnchannels(1), w.setsampwidth(2), w.setframerate(16000)
                w.writeframes(np.full(n, k, "<i2").tobytes())
    with open(tmp_path / "spk2gen.pkl", "wb") as f:
        pickle.dump({"p300": "F", "p225": "M"}, f)
    fake = _RecordingFrontEnd()
    monkeypatch.setattr(script, "default_frontend", lambda device=None: fake)
    script.make_spect_f0(str(root), str(tmp_path / "spmel"), str(tmp_path / "raptf0"),
                         str(tmp_path / "spk2gen.pkl"), verbose=False)
    assert len(fake.calls) == 1
    c = fake.calls[0]
    assert c["dtype"] == np.int16 and c["n"] == 5
    assert c["seed"] == [225, 225, 300, 300, 300]
    assert c["lo"] == [50.0, 50.0, 100.0, 100.0, 100.0] and c["hi"] == [250.0, 250.0, 600.0, 600.0, 600.0]
    # sorted file order within a speaker is k = 1, 0 (p225) and k = 2, 1, 0 (p300); 2560 and 5120 get the append
    assert c["skip"] == [0, 2561, 0, 7000, 11000]
    for spk, ls in lens.items():
        for k, n in enumerate(ls):
            stem = "%s_%03d.npy" % (spk, 9 - k)
            S = np.load(tmp_path / "spmel" / spk / stem, allow_pickle=False)
            f0 = np.load(tmp_path / "raptf0" / spk / stem, allow_pickle=False)
            T = (n + (n % 256 == 0) + 256) // 256
            assert S.shape == (T, 80) and f0.shape == (T,) and S.dtype == np.float32 and f0.dtype == np.float32
            with open(tmp_path / "spmel" / spk / stem, "rb") as fh:
                assert fh.read(8) == b"\x93NUMPY\x01\x00"                                     # NPY v1.0
    # each file got the rows of its own batch position
    assert np.load(tmp_path / "raptf0" / "p300" / "p300_009.npy")[0] == 4.0      # k = 0 sorts last
    assert np.load(tmp_path / "raptf0" / "p225" / "p225_008.npy")[0] == 0.0      # k = 1 sorts first

    with open(tmp_path / "spk2gen.pkl", "wb") as f:
        pickle.dump({"p300": "F", "p225": "X"}, f)
    with pytest.raises(ValueError):
        script.make_spect_f0(str(root), str(tmp_path / "o1"), str(tmp_path / "o2"), str(tmp_path / "spk2gen.pkl"),
                             verbose=False)


def test_script_form_batches_are_bounded():
    """extract_speakers consumes its speakers lazily and flushes every max_utts_per_call utterances; a
    speaker split over two calls continues its stream where the first call stopped."""
    from speechsplit_b200.make_spect_f0 import extract_speakers

    fake = _RecordingFrontEnd()
    pulled = []

    def speakers():
        for s in ("p226", "p227"):
            pulled.append(s)
            yield s, "M", [np.zeros(1000, np.int16)] * 3

    it = extract_speakers(fake, speakers(), max_utts_per_call=2)
    first = next(it)
    assert first[0] == "p226" and first[1] == 0 and pulled == ["p226"]       # p227 not read yet
    rest = list(it)
    assert [(r[0], r[1]) for r in [first] + rest] == [("p226", 0), ("p226", 1), ("p226", 2),
                                                     ("p227", 0), ("p227", 1), ("p227", 2)]
    assert [c["n"] for c in fake.calls] == [2, 2, 2]
    assert [c["skip"] for c in fake.calls] == [[0, 1000], [2000, 0], [1000, 2000]]
    assert [c["seed"] for c in fake.calls] == [[226, 226], [226, 227], [227, 227]]


# ---- data_loader / make_metadata mirror (host logic; the crop / clip / pad arithmetic is -m gpu) ----------
class _NumpyCollateFrontEnd:
    """Test double for FrontEnd in the loader tests: `_dev` keeps arrays on the host and `collate` is the
    reference's own per-item numpy code (data_loader.py:108-116), so that the HOST side of the mirror - the
    draws, the indices handed to the kernel, the batch layout - can be compared with the reference loop."""

    def _dev(self, a, dtype=None):
        return torch.from_numpy(np.ascontiguousarray(a))

    def collate(self, mel, f0, off, utt, left, len_crop, max_len_pad, want_onehot=False):
        mel, f0 = mel.numpy(), f0.numpy()
        a_all, c_all = [], []
        for u, l, n in zip(utt, left, len_crop):
            a = np.clip(mel[off[u] + l: off[u] + l + n], 0, 1)
            c = f0[off[u] + l: off[u] + l + n]
            a_all.append(np.pad(a, ((0, max_len_pad - n), (0, 0)), "constant"))
            c_all.append(np.pad(c[:, None], ((0, max_len_pad - n), (0, 0)), "constant", constant_values=-1e10))
        return torch.from_numpy(np.stack(a_all)), torch.from_numpy(np.stack(c_all)), None, None


def _feature_tree(tmp_path, spec, seed=3):
    rng = np.random.default_rng(seed)
    feats = {}
    for spk, Ts in spec.items():
        (tmp_path / "spmel" / spk).mkdir(parents=True)
        (tmp_path / "raptf0" / spk).mkdir(parents=True)
        for k, T in enumerate(Ts):
            S = (rng.random((T, 80)) * 1.3 - 0.15).astype(np.float32)
            f0 = rng.random(T).astype(np.float32)
            f0[rng.random(T) < 0.3] = -1e10
            name = "%s_%03d.npy" % (spk, 9 - k)
            np.save(tmp_path / "spmel" / spk / name, S, allow_pickle=False)
            np.save(tmp_path / "raptf0" / spk / name, f0, allow_pickle=False)
            feats[(spk, name)] = (S, f0)
    return feats


def _reference_collate(batch, min_len_seq, max_len_seq, max_len_pad):
    """data_loader.py:101-128 through the oracle restatement (pinned by tests/golden/collate.npz)."""
    from oracle import collate_ref
    return tuple(torch.from_numpy(a) for a in collate_ref.collate(batch, min_len_seq, max_len_seq, max_len_pad))


def _golden_tree(tmp_path, golden_dir):
    """The items of tests/golden/collate.npz as a spmel / raptf0 tree: item k is the only file of speaker
    number k (p226 is the one with the embedding at index 1), so dataset index == golden item index."""
    g = np.load(os.path.join(golden_dir, "collate.npz"))
    names = ["p225", "p226", "p227", "p228", "p229"]
    for k, spk in enumerate(names):
        (tmp_path / "spmel" / spk).mkdir(parents=True)
        (tmp_path / "raptf0" / spk).mkdir(parents=True)
        np.save(tmp_path / "spmel" / spk / (spk + "_001.npy"), g["S%d" % k], allow_pickle=False)
        np.save(tmp_path / "raptf0" / spk / (spk + "_001.npy"), g["f0%d" % k], allow_pickle=False)
    return g


def test_make_metadata_layout(tmp_path):
    """make_metadata.py:10-33: sorted speakers, [name, one-hot(82) f32, sorted 'spk/file.npy'...], train.pkl."""
    import pickle

    from speechsplit_b200.data_loader import make_metadata
    _feature_tree(tmp_path, {"p300": [150, 140], "p226": [200]})
    meta = make_metadata(str(tmp_path / "spmel"), verbose=False)
    with open(tmp_path / "spmel" / "train.pkl", "rb") as f:
        disk = pickle.load(f)
    assert [m[0] for m in meta] == ["p226", "p300"] == [m[0] for m in disk]
    assert meta[0][2:] == [os.path.join("p226", "p226_009.npy")]
    assert meta[1][2:] == [os.path.join("p300", "p300_008.npy"), os.path.join("p300", "p300_009.npy")]
    for m, hot in zip(meta, (1, 7)):
        assert m[1].dtype == np.float32 and m[1].shape == (82,) and m[1].sum() == 1.0 and m[1][hot] == 1.0
    assert np.array_equal(disk[1][1], meta[1][1])


def test_loader_mirror_host_logic(tmp_path):
    """Utterances / MyCollator / MultiSampler / get_loader against the reference's loop on the same seed:
    same items, same numpy draws in the same order, same batch layout and dtypes (solver.py:142)."""
    from types import SimpleNamespace

    from speechsplit_b200.data_loader import MultiSampler, Utterances, get_loader, make_metadata
    feats = _feature_tree(tmp_path, {"p225": [150, 400], "p226": [200], "p227": [135, 140]})
    make_metadata(str(tmp_path / "spmel"), verbose=False)
    hp = SimpleNamespace(root_dir=str(tmp_path / "spmel"), feat_dir=str(tmp_path / "raptf0"), mode="train",
                         batch_size=4, shuffle=False, num_workers=0, samplier=3, min_len_seq=64, max_len_seq=128,
                         max_len_pad=192)
    loader = get_loader(hp, frontend=_NumpyCollateFrontEnd())
    ds = loader.dataset
    # an item is a speaker with the FIRST file of its sorted list (data_loader.py:62-63)
    assert len(ds) == 3 and len(loader) == (3 * 3) // 4
    first = {"p225": "p225_008.npy", "p226": "p226_009.npy", "p227": "p227_008.npy"}
    for i, spk in enumerate(sorted(first)):
        melsp, emb, f0 = ds[i]
        assert np.array_equal(melsp, feats[(spk, first[spk])][0]) and np.array_equal(f0, feats[(spk, first[spk])][1])
        assert emb[1 if spk == "p226" else 7] == 1.0
    order = list(MultiSampler(3, 3).gen_sample_array().numpy())
    assert order == [0, 1, 2] * 3

    np.random.seed(11)
    ours = list(loader)
    np.random.seed(11)
    for b, got in enumerate(ours):
        items = [tuple(ds[i]) for i in order[4 * b: 4 * b + 4]]
        want = _reference_collate(items, 64, 128, 192)
        assert len(got) == 4
        for g, w, shape, dt in zip(got, want, [(4, 192, 80), (4, 82), (4, 192, 1), (4,)],
                                   [torch.float32, torch.float32, torch.float32, torch.int64]):
            assert tuple(g.shape) == shape and g.dtype == dt and w.dtype == dt
            assert torch.equal(g, w)
    assert len(ours) == 2

    with pytest.raises(ValueError):
        Utterances(hp.root_dir, hp.feat_dir, "valid")                    # data_loader.py:48-49
    # 'test' mode keeps frames [:split] with split = 0, i.e. nothing (data_loader.py:67-69)
    assert Utterances(hp.root_dir, hp.feat_dir, "test", frontend=_NumpyCollateFrontEnd())[0][0].shape == (0, 80)


def test_collator_refuses_short_utterance(tmp_path):
    """An utterance no longer than the drawn crop makes numpy's randint(0, <= 0) raise, as in the reference."""
    from types import SimpleNamespace

    from speechsplit_b200.data_loader import get_loader, make_metadata
    _feature_tree(tmp_path, {"p225": [60]})
    make_metadata(str(tmp_path / "spmel"), verbose=False)
    hp = SimpleNamespace(root_dir=str(tmp_path / "spmel"), feat_dir=str(tmp_path / "raptf0"), mode="train",
                         batch_size=1, shuffle=False, num_workers=0, samplier=1, min_len_seq=64, max_len_seq=128,
                         max_len_pad=192)
    with pytest.raises(ValueError):
        next(iter(get_loader(hp, frontend=_NumpyCollateFrontEnd())))


def test_make_metadata_matches_reference(tmp_path, golden_dir):
    """make_metadata against the train.pkl the reference's own make_metadata.py wrote for the same tree."""
    from speechsplit_b200.data_loader import make_metadata
    g = np.load(os.path.join(golden_dir, "collate.npz"))
    for e in g["meta_tree"]:
        spk, f = str(e).split("/")
        (tmp_path / "spmel" / spk).mkdir(parents=True, exist_ok=True)
        np.save(tmp_path / "spmel" / spk / f, np.zeros((1, 80), np.float32))
    meta = make_metadata(str(tmp_path / "spmel"), verbose=False)
    assert [m[0] for m in meta] == [str(x) for x in g["meta_speakers"]]
    assert np.array_equal(np.stack([m[1] for m in meta]), g["meta_emb"]) and meta[0][1].dtype == np.float32
    assert ["|".join(m[2:]) for m in meta] == [str(x) for x in g["meta_files"]]


def test_collator_draws_match_reference(tmp_path, golden_dir):
    """The crops MyCollator hands to the kernel are the ones the reference's own collator took: same
    lengths (len_org of the golden batch) and same positions (the golden rows are found at `left`)."""
    from types import SimpleNamespace

    from speechsplit_b200.data_loader import MyCollator, Utterances, make_metadata
    g = _golden_tree(tmp_path, golden_dir)
    make_metadata(str(tmp_path / "spmel"), verbose=False)
    ds = Utterances(str(tmp_path / "spmel"), str(tmp_path / "raptf0"), "train", frontend=_NumpyCollateFrontEnd())
    col = MyCollator(SimpleNamespace(min_len_seq=64, max_len_seq=128, max_len_pad=192), ds)
    np.random.seed(int(g["seed"]))
    utt, left, len_crop = col.draw([ds[int(i)] for i in g["order"]])
    assert np.array_equal(utt, g["order"]) and np.array_equal(len_crop, g["len_org"])
    for b, (u, l, n) in enumerate(zip(utt, left, len_crop)):
        assert np.array_equal(g["pitch"][b, :n, 0], g["f0%d" % u][l:l + n])
    np.random.seed(int(g["seed"]))
    melsp, spk_emb, pitch, len_org = col([ds[int(i)] for i in g["order"]])
    assert torch.equal(spk_emb, torch.from_numpy(g["spk_emb"])) and torch.equal(len_org, torch.from_numpy(g["len_org"]))


def test_collator_batched_draws(tmp_path):
    """draws='batched': same distributions as data_loader.py:104-105 (length uniform on [min, max], left
    uniform on [0, T - length)), two numpy calls per batch; crops always inside the utterance."""
    from types import SimpleNamespace

    from speechsplit_b200.data_loader import MyCollator, Utterances, make_metadata
    _feature_tree(tmp_path, {"p225": [129], "p226": [300]})
    make_metadata(str(tmp_path / "spmel"), verbose=False)
    ds = Utterances(str(tmp_path / "spmel"), str(tmp_path / "raptf0"), "train", frontend=_NumpyCollateFrontEnd())
    hp = SimpleNamespace(min_len_seq=64, max_len_seq=128, max_len_pad=192)
    col = MyCollator(hp, ds, draws="batched")
    np.random.seed(2)
    lens, lefts = [], []
    for _ in range(400):
        utt, left, len_crop = col.draw([ds[0], ds[1], ds[1], ds[0]])
        assert list(utt) == [0, 1, 1, 0] and left.dtype == np.int32 and len_crop.dtype == np.int64
        assert np.all(left >= 0) and np.all(left + len_crop < np.array([129, 300, 300, 129]))
        lens.append(len_crop)
        lefts.append(left)
    lens, lefts = np.concatenate(lens), np.stack(lefts)
    assert lens.min() == 64 and lens.max() == 128 and abs(lens.mean() - 96.0) < 2.0
    assert lefts[:, 1].max() > 200 and lefts[:, 0].max() <= 64
    melsp, spk_emb, pitch, len_org = col([ds[1], ds[0]])
    assert tuple(melsp.shape) == (2, 192, 80) and tuple(pitch.shape) == (2, 192, 1) and len_org.dtype == torch.int64
    with pytest.raises(ValueError):
        MyCollator(hp, ds, draws="fast")
    _feature_tree(tmp_path / "short", {"p225": [100]})
    make_metadata(str(tmp_path / "short" / "spmel"), verbose=False)
    short = Utterances(str(tmp_path / "short" / "spmel"), str(tmp_path / "short" / "raptf0"), "train",
                       frontend=_NumpyCollateFrontEnd())
    with pytest.raises(ValueError):
        for _ in range(200):                                     # a crop of >= 100 frames is drawn soon enough
            MyCollator(hp, short, draws="batched").draw([short[0]])


def test_read_wav_sample_formats(tmp_path):
    """read_wav without soundfile: 16-bit PCM stays int16 (x = v / 32768 exactly, what sf.read returns);
    8 / 24 / 32-bit PCM and IEEE-float files come back as the float64 soundfile would give; stereo is refused."""
    import wave

    from scipy.io import wavfile

    from speechsplit_b200.make_spect_f0 import read_wav
    rng = np.random.default_rng(4)
    v16 = rng.integers(-32768, 32768, 1000).astype(np.int16)
    wavfile.write(tmp_path / "a16.wav", 16000, v16)
    x, fs = read_wav(str(tmp_path / "a16.wav"))
    assert fs == 16000 and x.dtype == np.int16 and np.array_equal(x, v16)
    v32 = rng.integers(-2**31, 2**31, 1000).astype(np.int32)
    wavfile.write(tmp_path / "a32.wav", 16000, v32)
    x, _ = read_wav(str(tmp_path / "a32.wav"))
    assert x.dtype == np.float64 and np.array_equal(x, v32 / 2147483648.0)
    v8 = rng.integers(0, 256, 1000).astype(np.uint8)
    wavfile.write(tmp_path / "a8.wav", 16000, v8)
    x, _ = read_wav(str(tmp_path / "a8.wav"))
    assert x.dtype == np.float64 and np.array_equal(x, (v8.astype(np.float64) - 128) / 128)
    vf = (rng.random(1000) - 0.5).astype(np.float32)
    wavfile.write(tmp_path / "af.wav", 16000, vf)
    x, _ = read_wav(str(tmp_path / "af.wav"))
    assert x.dtype == np.float64 and np.array_equal(x, vf.astype(np.float64))
    v24 = rng.integers(-2**23, 2**23, 1000)
    with wave.open(str(tmp_path / "a24.wav"), "wb") as w:
        w.setnchannels(1), w.setsampwidth(3), w.setframerate(16000)
        w.writeframes(b"".join(int(t).to_bytes(3, "little", signed=True) for t in v24))
    x, _ = read_wav(str(tmp_path / "a24.wav"))
    assert x.dtype == np.float64 and np.array_equal(x, v24 / 8388608.0)
    wavfile.write(tmp_path / "st.wav", 16000, np.stack([v16, v16], 1))
    with pytest.raises(ValueError):
        read_wav(str(tmp_path / "st.wav"))


def _riff(chunks):
    body = b"WAVE" + b"".join(cid + len(data).to_bytes(4, "little") + data + (b"\0" if len(data) & 1 else b"")
                              for cid, data in chunks)
    return b"RIFF" + len(body).to_bytes(4, "little") + body


def test_read_wav_pcm16_fast_path(tmp_path):
    """read_wav_pcm16 (one read + a walk over the chunk headers) against the stdlib reader, on the layouts real files
    have: extra chunks before and after the data, odd-sized chunks and their pad byte, WAVE_FORMAT_EXTENSIBLE, a data
    chunk that claims more than the file holds; everything that is not mono 16-bit PCM is left to the general reader."""
    import struct
    import wave

    from speechsplit_b200.make_spect_f0 import read_wav, read_wav_pcm16
    rng = np.random.default_rng(11)
    v = rng.integers(-32768, 32768, 777).astype("<i2")
    fmt = struct.pack("<HHIIHH", 1, 1, 16000, 32000, 2, 16)
    plain = tmp_path / "plain.wav"
    with wave.open(str(plain), "wb") as w:
        w.setnchannels(1), w.setsampwidth(2), w.setframerate(16000)
        w.writeframes(v.tobytes())
    x, fs = read_wav_pcm16(str(plain))
    assert fs == 16000 and x.dtype == np.int16 and np.array_equal(x, v) and not x.flags.writeable
    # LIST chunk of odd size in front (pad byte), a trailing chunk behind the data
    p = tmp_path / "chunks.wav"
    p.write_bytes(_riff([(b"LIST", b"INFOabc"), (b"fmt ", fmt), (b"fact", b"\1\0\0\0"), (b"data", v.tobytes()), (b"cue ", b"xy")]))
    x, fs = read_wav_pcm16(str(p))
    assert fs == 16000 and np.array_equal(x, v)
    with wave.open(str(p), "rb") as w:
        assert np.array_equal(np.frombuffer(w.readframes(w.getnframes()), "<i2"), x)
    # extensible format with the PCM sub-format GUID
    ext = struct.pack("<HHIIHHHHIH", 0xFFFE, 1, 16000, 32000, 2, 16, 22, 16, 4, 1) + bytes.fromhex("000000001000800000aa00389b71")
    assert len(ext) == 40
    p = tmp_path / "ext.wav"
    p.write_bytes(_riff([(b"fmt ", ext), (b"data", v.tobytes())]))
    x, fs = read_wav_pcm16(str(p))
    assert fs == 16000 and np.array_equal(x, v)
    # truncated file: the data chunk claims 1554 bytes, 1001 are there -> 500 whole samples
    raw = _riff([(b"fmt ", fmt), (b"data", v.tobytes())])
    p = tmp_path / "trunc.wav"
    p.write_bytes(raw[:len(raw) - 553])
    x, _ = read_wav_pcm16(str(p))
    assert np.array_equal(x, v[:500])
    # an empty data chunk is an empty utterance, not an error here (ssfe_extract refuses it with its own message)
    p = tmp_path / "empty.wav"
    p.write_bytes(_riff([(b"fmt ", fmt), (b"data", b"")]))
    x, _ = read_wav_pcm16(str(p))
    assert x.shape == (0,)
    # not for the fast path: stereo, 8-bit, float, data before fmt, not RIFF at all, a short file
    for name, chunks in [("st", [(b"fmt ", struct.pack("<HHIIHH", 1, 2, 16000, 64000, 4, 16)), (b"data", v.tobytes()[:776 * 2])]),
                         ("u8", [(b"fmt ", struct.pack("<HHIIHH", 1, 1, 16000, 16000, 1, 8)), (b"data", b"\x80" * 10)]),
                         ("f32", [(b"fmt ", struct.pack("<HHIIHH", 3, 1, 16000, 64000, 4, 32)), (b"data", b"\0" * 16)]),
                         ("nofmt", [(b"data", v.tobytes())])]:
        q = tmp_path / (name + ".wav")
        q.write_bytes(_riff(chunks))
        assert read_wav_pcm16(str(q)) is None
    (tmp_path / "junk.wav").write_bytes(b"OggS" + b"\0" * 64)
    (tmp_path / "tiny.wav").write_bytes(b"RIFF")
    assert read_wav_pcm16(str(tmp_path / "junk.wav")) is None and read_wav_pcm16(str(tmp_path / "tiny.wav")) is None
    # ... and read_wav hands those to the general reader: stereo is refused there, 8-bit comes back as float64
    with pytest.raises(ValueError):
        read_wav(str(tmp_path / "st.wav"))
    x, _ = read_wav(str(tmp_path / "u8.wav"))
    assert x.dtype == np.float64 and np.array_equal(x, np.zeros(10))


def test_save_npy_is_np_save_byte_for_byte(tmp_path):
    """make_spect_f0.py:71-74 writes with np.save(..., allow_pickle=False); save_npy must leave the same bytes (NPY v1.0
    header included) for the shapes the script writes, for views, empty arrays and paths with or without '.npy'."""
    from speechsplit_b200.make_spect_f0 import save_npy
    rng = np.random.default_rng(5)
    big = rng.random((300, 80)).astype(np.float32)
    cases = [big[:188], big[17:18], big[:0], big[:, 3].copy(), big[:, 3], big[::2], rng.random(188).astype(np.float32),
             np.arange(7, dtype=np.int64), rng.random((4, 3, 2))]
    for i, a in enumerate(cases * 2):                      # twice: the second round takes the cached headers
        ref = tmp_path / ("ref%d.npy" % i)
        np.save(ref, a, allow_pickle=False)
        got = tmp_path / ("got%d" % i)
        save_npy(str(got) + (".npy" if i & 1 else ""), a)
        assert (tmp_path / ("got%d.npy" % i)).read_bytes() == ref.read_bytes(), i
        assert np.array_equal(np.load(tmp_path / ("got%d.npy" % i), allow_pickle=False), a)
    with pytest.raises(ValueError):
        save_npy(str(tmp_path / "o"), np.array([{}], dtype=object))


def test_script_form_io_threads_give_the_same_trees(tmp_path, monkeypatch):
    """The inline file handling (default) and the opt-in thread pool write identical trees."""
    import pickle
    import wave

    from speechsplit_b200 import make_spect_f0 as script
    root = tmp_path / "wavs"
    rng = np.random.default_rng(2)
    for spk in ("p226", "p231", "p240"):
        (root / spk).mkdir(parents=True)
        for k in range(7):
            with wave.open(str(root / spk / ("%s_%03d.wav" % (spk, k + 1))), "wb") as w:
                w.setnchannels(1), w.setsampwidth(2), w.setframerate(16000)
                w.writeframes(rng.integers(-3000, 3000, int(rng.integers(2000, 9000))).astype("<i2").tobytes())
    with open(tmp_path / "spk2gen.pkl", "wb") as f:
        pickle.dump({"p226": "M", "p231": "F", "p240": "F"}, f)
    trees = []
    for threads in (1, 4):
        fake = _RecordingFrontEnd()
        out = tmp_path / ("t%d" % threads)
        st = {}
        script.make_spect_f0(str(root), str(out / "spmel"), str(out / "raptf0"), str(tmp_path / "spk2gen.pkl"),
                             verbose=False, io_threads=threads, frontend=fake, stats=st)
        assert st["files"] == 21 and st["calls"] == 1 and len(fake.calls) == 1
        trees.append({os.path.relpath(os.path.join(d, f), out): open(os.path.join(d, f), "rb").read()
                      for d, _, fs in os.walk(out) for f in fs})
    assert len(trees[0]) == 42 and trees[0] == trees[1]
