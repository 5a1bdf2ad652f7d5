"""Property tests (hypothesis) of the host-side logic around the kernels - CPU only - and, on a GPU, of the
quantiser against the oracle on arbitrary inputs (the reference has no tests; these pin the
contracts the drop-in relies on for inputs nobody wrote down)."""
import numpy as np
import pytest
from hypothesis import given, settings
from hypothesis import strategies as st

from oracle import ref_pipeline as rp
from speechsplit_b200 import _lib
from speechsplit_b200.sharding import contiguous_shards, dither_skips, fixed_length, lpt_shards

lengths_st = st.lists(st.integers(min_value=1, max_value=200000), min_size=1, max_size=60)


@settings(max_examples=200, deadline=None)
@given(lengths_st)
def test_plan_offsets_match_the_reference_arithmetic(lengths):
    """make_spect_f0.py:52-53 (append when L % 256 == 0), utils.py:22-23 (frames) and :69 (same count as RAPT's
    ceil(L' / 256)) for arbitrary lengths, through the C ABI's planner."""
    import ctypes
    lib = _lib.load()
    off = np.concatenate([[0], np.cumsum(lengths)]).astype(np.int64)
    n = len(lengths)
    fix, fr = np.empty(n + 1, np.int64), np.empty(n + 1, np.int64)
    p = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))
    assert lib.ssfe_plan_offsets(p(off), n, p(fix), p(fr)) == 0
    for i, L in enumerate(lengths):
        Lf = L + 1 if L % 256 == 0 else L
        assert fix[i + 1] - fix[i] == Lf == fixed_length(L) == lib.ssfe_fixed_length(L)
        frames = (Lf + 1024 - 768) // 256                     # pySTFT: (len + 2*512 - noverlap) // hop
        assert fr[i + 1] - fr[i] == frames == -(-Lf // 256) == lib.ssfe_num_frames(L)
    assert fix[0] == 0 and fr[0] == 0


@settings(max_examples=100, deadline=None)
@given(lengths_st, st.integers(min_value=1, max_value=8))
def test_contiguous_shards_partition_the_corpus_in_order(lengths, world):
    shards = contiguous_shards(lengths, world)
    assert len(shards) == world
    flat = np.concatenate([np.asarray(s, np.int64) for s in shards]) if shards else np.zeros(0, np.int64)
    assert np.array_equal(flat, np.arange(len(lengths)))          # every utterance exactly once, corpus order kept
    total, worst = sum(lengths), max(lengths)
    loads = [int(sum(lengths[i] for i in s)) for s in shards]
    assert sum(loads) == total
    assert max(loads) <= total / world + worst + 1e-9             # balanced to within one utterance


@settings(max_examples=100, deadline=None)
@given(lengths_st, st.integers(min_value=1, max_value=8))
def test_lpt_shards_cover_everything(lengths, world):
    shards = lpt_shards(lengths, world)
    flat = np.sort(np.concatenate([np.asarray(s, np.int64) for s in shards]))
    assert np.array_equal(flat, np.arange(len(lengths)))
    for s in shards:
        assert np.all(np.diff(np.asarray(s)) > 0)                 # corpus order inside a shard (sorted stream requests)


@settings(max_examples=100, deadline=None)
@given(st.lists(st.tuples(st.sampled_from(["p225", "p226", "p300"]), st.integers(min_value=1, max_value=100000)),
                min_size=1, max_size=40))
def test_dither_skips_are_the_stream_positions_of_the_reference_loop(items):
    """make_spect_f0.py:47-55: one RandomState per speaker, consumed with the post-append length of every file."""
    spk = [s for s, _ in items]
    lens = [n for _, n in items]
    skips = dither_skips(spk, lens)
    pos = {}
    for s, n, k in zip(spk, lens, skips):
        assert int(k) == pos.get(s, 0)
        pos[s] = pos.get(s, 0) + (n + 1 if n % 256 == 0 else n)


f0_values = st.one_of(st.just(-1e10), st.just(0.0), st.floats(min_value=0.0, max_value=1.0, allow_nan=False),
                      st.integers(min_value=0, max_value=510).map(lambda k: k / 510.0))     # incl. the half-way ties


@settings(max_examples=200, deadline=None)
@given(st.lists(f0_values, min_size=1, max_size=300))
def test_oracle_quantiser_contract(vals):
    """utils.py:46-58: unvoiced (<= 0) -> bin 0, voiced -> round-half-even(x * 255) + 1, one-hot rows."""
    x = np.asarray(vals, np.float64)
    enc, idx = rp.quantize_f0_numpy(x)
    assert enc.shape == (len(x), 257) and enc.dtype == np.float32 and idx.dtype == np.int64
    for v, b, row in zip(x, idx, enc):
        want = 0 if v <= 0 else int(np.round(v * 255)) + 1
        assert b == want and row[b] == 1.0 and row.sum() == 1.0
    assert np.array_equal(x, np.asarray(vals, np.float64))        # the input is not mutated (.copy() at :49)


@pytest.mark.gpu
@settings(max_examples=25, deadline=None)
@given(st.lists(f0_values, min_size=1, max_size=2000), st.sampled_from([np.float32, np.float64]))
def test_gpu_quantiser_equals_oracle_on_arbitrary_input(vals, dtype):
    from speechsplit_b200 import utils
    x = np.asarray(vals, dtype)
    enc, idx = utils.quantize_f0_numpy(x)
    renc, ridx = rp.quantize_f0_numpy(x)
    assert np.array_equal(idx, ridx) and np.array_equal(enc, renc)


_chunk_st = st.tuples(st.sampled_from([b"LIST", b"fact", b"cue ", b"bext", b"JUNK", b"id3 "]), st.binary(min_size=0, max_size=37))


@settings(max_examples=150, deadline=None)
@given(st.lists(_chunk_st, max_size=3), st.lists(_chunk_st, max_size=2), st.lists(_chunk_st, max_size=2),
       st.binary(min_size=0, max_size=401), st.sampled_from([8000, 16000, 22050, 48000]))
def test_read_wav_pcm16_agrees_with_the_stdlib_reader(before, between, after, payload, fs):
    """Arbitrary RIFF layouts around a mono 16-bit PCM payload (extra chunks of odd and even size before the format
    chunk, between it and the data, and behind it; odd payload sizes): read_wav_pcm16 returns what ``wave`` returns."""
    import io
    import os
    import struct
    import tempfile
    import wave

    from speechsplit_b200.make_spect_f0 import read_wav_pcm16
    fmt = struct.pack("<HHIIHH", 1, 1, fs, 2 * fs, 2, 16)
    chunks = list(before) + [(b"fmt ", fmt)] + list(between) + [(b"data", payload)] + list(after)
    body = b"WAVE" + b"".join(c + len(d).to_bytes(4, "little") + d + (b"\0" if len(d) & 1 else b"") for c, d in chunks)
    blob = b"RIFF" + len(body).to_bytes(4, "little") + body
    with wave.open(io.BytesIO(blob), "rb") as w:
        ref = np.frombuffer(w.readframes(w.getnframes()), dtype="<i2")
        ref_fs = w.getframerate()
    fd, path = tempfile.mkstemp(suffix=".wav")
    try:
        with os.fdopen(fd, "wb") as fh:
            fh.write(blob)
        got = read_wav_pcm16(path)
    finally:
        os.unlink(path)
    assert got is not None and got[1] == ref_fs == fs
    assert got[0].dtype == np.int16 and np.array_equal(got[0], ref)
