"""Function / CLI form of the reference script ``make_spect_f0.py`` on the GPU.

    python -m speechsplit_b200.make_spect_f0 [--root assets/wavs] [--out assets/spmel]
                                             [--out-f0 assets/raptf0] [--spk2gen assets/spk2gen.pkl]
                                             [--device N] [--io-threads N]

Same inputs and outputs as the reference (make_spect_f0.py:19-25,71-74): ``<root>/<spk>/*.wav``
at 16 kHz mono, a pickled dict speaker -> 'M' / 'F', speaker directories named ``p<int>``;
writes ``<out>/<spk>/<utt>.npy`` (T, 80) float32 and ``<out-f0>/<spk>/<utt>.npy`` (T,) float32
with ``np.save(..., allow_pickle=False)``.  The per-utterance loop body (:50-67) runs as one
``ssfe_extract`` call per group of speakers; file I/O stays on the host (mono 16-bit PCM through
``read_wav_pcm16``, anything else through the reference's ``soundfile`` when importable, else stdlib
``wave`` / scipy; NPY v1.0 through ``save_npy``).
"""
import argparse
import io
import os
import pickle
import struct
import time
import wave
from concurrent.futures import ThreadPoolExecutor

import numpy as np
from numpy.lib import format as npy_format

from .frontend import GENDER_RANGE, default_frontend
from .sharding import fixed_length


def read_wav_pcm16(path):
    """Fast path of read_wav: a RIFF/WAVE file holding mono 16-bit PCM (format tag 1, or WAVE_FORMAT_EXTENSIBLE
    with the PCM sub-format) -> (read-only int16 array, fs); None for anything else.  One read of the whole file
    and a walk over its chunk headers: 32 us per 3 s file against 116 us through the stdlib ``wave`` module
    (profiles/microbench/wav_npy_io.py) - in the script form the WAV reads were 1.0 s of a 1.4 s run."""
    with open(path, "rb") as fh:
        b = fh.read()
    nb = len(b)
    if nb < 12 or b[:4] != b"RIFF" or b[8:12] != b"WAVE":
        return None
    pos, fmt = 12, None
    while pos + 8 <= nb:
        cid = b[pos:pos + 4]
        size = int.from_bytes(b[pos + 4:pos + 8], "little")
        body = pos + 8
        if cid == b"fmt ":
            if size < 16 or body + 16 > nb:
                return None
            tag, nch, fs, _, align, bits = struct.unpack_from("<HHIIHH", b, body)
            if tag == 0xFFFE and size >= 40 and body + 26 <= nb:       # extensible: the sub-format GUID starts with the tag
                tag = struct.unpack_from("<H", b, body + 24)[0]
            if tag != 1 or nch != 1 or bits != 16 or align != 2:
                return None
            fmt = fs
        elif cid == b"data":
            if fmt is None:
                return None
            n = min(size, nb - body) // 2                              # whole frames only, like wave.readframes
            return np.frombuffer(b, dtype="<i2", count=n, offset=body), fmt
        pos = body + size + (size & 1)                                 # chunks are word aligned
    return None


_NPY_HEADERS = {}


def save_npy(path, arr):
    """``np.save(path, arr, allow_pickle=False)`` for a C-contiguous numeric array, byte for byte (NPY v1.0, same header
    padding - the header comes from numpy's own ``write_array_header_1_0``, cached per (shape, dtype)), without
    np.save's per-call checks: 25 us per file against 48 (profiles/microbench/wav_npy_io.py).  make_spect_f0.py:71-74."""
    arr = np.ascontiguousarray(arr)
    if arr.dtype.hasobject:
        raise ValueError("save_npy: object arrays are not allowed (allow_pickle=False)")
    key = (arr.shape, arr.dtype.str)
    head = _NPY_HEADERS.get(key)
    if head is None:
        bio = io.BytesIO()
        npy_format.write_array_header_1_0(bio, npy_format.header_data_from_array_1_0(arr))
        head = bio.getvalue()
        if len(_NPY_HEADERS) < 65536:
            _NPY_HEADERS[key] = head
    if not path.endswith(".npy"):
        path = path + ".npy"                                           # np.save appends the extension
    with open(path, "wb") as fh:
        fh.write(head)
        fh.write(arr.data)


def read_wav(path):
    """-> (int16 or float64 array, fs).  16-bit PCM stays int16 (x = v/32768 exactly, like sf.read);
    any other sample format goes through soundfile as float64, as at make_spect_f0.py:50."""
    got = read_wav_pcm16(path)
    if got is not None:
        return got
    try:
        import soundfile as sf
    except ImportError:
        sf = None
    if sf is not None:
        if sf.info(path).subtype == "PCM_16":
            return sf.read(path, dtype="int16")
        return sf.read(path)
    try:
        with wave.open(path, "rb") as w:
            fs, nch, sw, n = w.getframerate(), w.getnchannels(), w.getsampwidth(), w.getnframes()
            raw = w.readframes(n) if (sw == 2 and nch == 1) else None
    except wave.Error:                      # e.g. IEEE-float WAV (format tag 3), which the stdlib does not read
        raw, nch = None, 1
    if raw is not None:
        return np.frombuffer(raw, dtype="<i2"), fs
    if nch != 1:
        raise ValueError("%s: %d channels; the pipeline takes mono files (make_spect_f0.py:50-52)" % (path, nch))
    # other sample formats: scipy's reader, scaled to [-1, 1) the way soundfile's float64 read does
    from scipy.io import wavfile
    fs, data = wavfile.read(path)
    if data.ndim != 1:
        raise ValueError("%s: the pipeline takes mono files (make_spect_f0.py:50-52)" % path)
    if data.dtype == np.int16:
        return data, fs
    if data.dtype == np.uint8:
        return (data.astype(np.float64) - 128.0) / 128.0, fs
    if data.dtype == np.int32:              # 32-bit PCM, and 24-bit PCM left-justified in 32 bits
        return data.astype(np.float64) / 2147483648.0, fs
    if data.dtype in (np.float32, np.float64):
        return data.astype(np.float64), fs
    raise ValueError("%s: unsupported WAV sample format %s" % (path, data.dtype))


def extract_speakers(fe, speakers, max_utts_per_call=4096, stats=None):
    """speakers: iterable of (spk_name, gender, [arrays in sorted file order]), consumed lazily.  Yields
    (spk_name, file_index, S (T,80) f32, f0_norm (T,) f32) in the reference's loop order.
    stats: optional dict, gets the seconds spent in the GPU calls ('extract_s') and in batching ('pack_s')."""
    batch, meta = [], []

    def flush():
        if not batch:
            return
        t0 = time.perf_counter()
        if all(a.dtype == np.int16 for a in batch):
            x = np.concatenate(batch)
        else:       # mixed sample formats: everything as the float64 sf.read would have returned
            x = np.concatenate([a / 32768.0 if a.dtype == np.int16 else a.astype(np.float64) for a in batch])
        off = np.concatenate([[0], np.cumsum([len(a) for a in batch])]).astype(np.int64)
        lo = [GENDER_RANGE[m[1]][0] for m in meta]
        hi = [GENDER_RANGE[m[1]][1] for m in meta]
        t1 = time.perf_counter()
        res = fe.extract_host(x, off, lo, hi, [m[2] for m in meta], [m[3] for m in meta], want_bins=False)
        if stats is not None:
            stats["pack_s"] = stats.get("pack_s", 0.0) + (t1 - t0)
            stats["extract_s"] = stats.get("extract_s", 0.0) + (time.perf_counter() - t1)
            stats["calls"] = stats.get("calls", 0) + 1
        fo = res["frame_offsets"]
        for i, m in enumerate(meta):
            yield m[0], m[4], res["mel"][fo[i]:fo[i + 1]], res["f0_norm"][fo[i]:fo[i + 1]]
        batch.clear()
        meta.clear()

    for spk, gender, utts in speakers:
        if gender not in GENDER_RANGE:
            raise ValueError                                   # make_spect_f0.py:45
        seed = int(spk[1:])                                    # :47
        skip = 0
        for k, a in enumerate(utts):
            batch.append(np.asarray(a))
            meta.append((spk, gender, seed, skip, k))
            skip += fixed_length(len(a))                       # the stream advances by the post-append length (:53,55)
            if len(batch) >= max_utts_per_call:
                yield from flush()
    yield from flush()


def make_spect_f0(root_dir="assets/wavs", target_dir="assets/spmel", target_dir_f0="assets/raptf0",
                  spk2gen_path="assets/spk2gen.pkl", device=None, verbose=True, stats=None, io_threads=1, frontend=None):
    """stats: optional dict that receives where the wall time went (read_s, pack_s, extract_s, write_s, files).
    frontend: the FrontEnd to use (default: the process-wide one on ``device``; the CPU tests of the file handling pass a stub).
    io_threads: > 1 sends WAV reads and NPY writes through a thread pool (for slow file systems; see below)."""
    spk2gen = pickle.load(open(spk2gen_path, "rb"))            # :19
    # io_threads <= 1: reads and writes inline.  A file costs 30-70 us to read and 25-40 us to write, less than a
    # hand-over to a worker thread and back under the GIL: through an 8-thread pool the same tree took 2.07 s against
    # 1.0 s inline (profiles/microbench/wav_npy_io.py, stub front end), so the pool is opt-in for slow file systems.
    n_threads = max(1, int(io_threads))
    pool = ThreadPoolExecutor(max_workers=n_threads) if n_threads > 1 else None
    pending = []
    fe = frontend if frontend is not None else default_frontend(device)
    dir_name, subdirs, _ = next(os.walk(root_dir))             # :28
    if verbose:
        print("Found directory: %s" % dir_name)
    names = {}

    def speakers():
        # read lazily: extract_speakers holds at most one call's worth of utterances plus one speaker
        for subdir in sorted(subdirs):                         # :31
            if verbose:
                print(subdir)
            os.makedirs(os.path.join(target_dir, subdir), exist_ok=True)
            os.makedirs(os.path.join(target_dir_f0, subdir), exist_ok=True)
            _, _, files = next(os.walk(os.path.join(dir_name, subdir)))
            files = sorted(files)                              # :48
            t0 = time.perf_counter()
            src = os.path.join(dir_name, subdir)
            if pool is not None:
                got = list(pool.map(lambda f: read_wav(os.path.join(src, f)), files))
            else:
                got = [read_wav(os.path.join(src, f)) for f in files]
            utts = []
            for x, fs in got:
                assert fs == 16000                             # :51
                utts.append(x)
            if stats is not None:
                stats["read_s"] = stats.get("read_s", 0.0) + (time.perf_counter() - t0)
            names[subdir] = files
            yield subdir, spk2gen[subdir], utts

    t0 = time.perf_counter()
    try:
        for spk, k, S, f0n in extract_speakers(fe, speakers(), stats=stats):
            stem = names[spk][k][:-4]
            t0 = time.perf_counter()
            # (pool: the rows are views of the call's result arrays, which stay alive until the writes are done)
            for path, arr in ((os.path.join(target_dir, spk, stem), S),                                      # :71-72
                              (os.path.join(target_dir_f0, spk, stem), f0n)):                                # :73-74
                arr = arr.astype(np.float32, copy=False)
                if pool is not None:
                    pending.append(pool.submit(save_npy, path, arr))
                else:
                    save_npy(path, arr)
            if stats is not None:
                stats["write_s"] = stats.get("write_s", 0.0) + (time.perf_counter() - t0)
                stats["files"] = stats.get("files", 0) + 1
        t0 = time.perf_counter()
        for f in pending:
            f.result()                                             # re-raises an I/O error of a worker
    finally:
        if pool is not None:
            pool.shutdown(wait=True)
    if stats is not None:
        stats["write_s"] = stats.get("write_s", 0.0) + (time.perf_counter() - t0)


def main():
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--root", default="assets/wavs")
    ap.add_argument("--out", default="assets/spmel")
    ap.add_argument("--out-f0", default="assets/raptf0")
    ap.add_argument("--spk2gen", default="assets/spk2gen.pkl")
    ap.add_argument("--device", type=int, default=None)
    ap.add_argument("--io-threads", type=int, default=1,
                    help="> 1: WAV reads and NPY writes through a thread pool (worth it on slow file systems only)")
    a = ap.parse_args()
    make_spect_f0(a.root, a.out, a.out_f0, a.spk2gen, a.device, io_threads=a.io_threads)


if __name__ == "__main__":
    main()
