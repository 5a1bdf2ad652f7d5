"""Host-side argument checks of speechsplit_b200.frontend (no GPU needed).

The C ABI takes plain pointers and sizes (include/ssfe.h), so everything the reference gets for free
from numpy indexing - offsets that leave the buffer, output buffers of the wrong shape or dtype - has to
be refused before a pointer crosses the boundary."""
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
