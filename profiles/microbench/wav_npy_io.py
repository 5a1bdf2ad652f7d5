"""Host-side file handling of the script form, timed WITHOUT a GPU.

    python profiles/microbench/wav_npy_io.py [n_speakers] [utts]

Builds the bench's WAV tree shape (109 speakers x 40 files of ~3 s, 16-bit PCM) on tmpfs and times
  * the WAV readers: stdlib ``wave`` (what read_wav used) against read_wav_pcm16 (one read + a chunk walk),
  * the NPY writers: np.save against save_npy (numpy's own v1.0 header, cached per shape),
  * the whole ``make_spect_f0`` flow with a STUB front end (zeros of the right shapes, no GPU work): what the
    script form costs around its ssfe_extract_host calls.
The stub is test scaffolding for the file handling only; nothing here measures or replaces the CUDA path.
"""
import os
import pickle
import shutil
import sys
import tempfile
import time
import wave

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from speechsplit_b200 import make_spect_f0 as msf      # noqa: E402
from speechsplit_b200.sharding import fixed_length     # noqa: E402


class StubFrontEnd:
    """extract_host with the real output shapes and no arithmetic."""

    def extract_host(self, x, off, lo, hi, seeds, skips, want_bins=True):
        fr = np.concatenate([[0], np.cumsum([(fixed_length(int(b - a)) + 256) // 256 for a, b in zip(off[:-1], off[1:])])])
        T = int(fr[-1])
        return dict(mel=np.zeros((T, 80), np.float32), f0_norm=np.zeros(T, np.float32), frame_offsets=fr)


def wave_read(path):
    with wave.open(path, "rb") as w:
        return np.frombuffer(w.readframes(w.getnframes()), dtype="<i2"), w.getframerate()


def best(fn, reps=3):
    t = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        t.append(time.perf_counter() - t0)
    return min(t)


def main():
    n_spk = int(sys.argv[1]) if len(sys.argv) > 1 else 109
    utts = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
    tmp = tempfile.mkdtemp(prefix="ssfe_io_", dir=base)
    rng = np.random.default_rng(0)
    try:
        paths, spk2gen = [], {}
        for s in range(n_spk):
            spk = "p%d" % (225 + s)
            spk2gen[spk] = "MF"[s & 1]
            os.makedirs(os.path.join(tmp, "wavs", spk))
            for u in range(utts):
                n = int(16000 * np.clip(rng.normal(3.0, 0.8), 1.0, 8.0))
                p = os.path.join(tmp, "wavs", spk, "%s_%03d.wav" % (spk, u + 1))
                with wave.open(p, "wb") as w:
                    w.setnchannels(1)
                    w.setsampwidth(2)
                    w.setframerate(16000)
                    w.writeframes((rng.standard_normal(n) * 3000).astype("<i2").tobytes())
                paths.append(p)
        with open(os.path.join(tmp, "spk2gen.pkl"), "wb") as fh:
            pickle.dump(spk2gen, fh)
        nf = len(paths)
        lens = []
        for p in paths:
            x, y = wave_read(p), msf.read_wav_pcm16(p)
            assert np.array_equal(x[0], y[0]) and x[1] == y[1]
            lens.append(len(x[0]))
        def drain(fn):          # results dropped at once: keeping 0.4 GB of arrays alive measures page faults, not readers
            for p in paths:
                fn(p)
        t_old = best(lambda: drain(wave_read))
        t_new = best(lambda: drain(msf.read_wav_pcm16))
        print("WAV read, %d files, one thread: wave module %.3f s (%.0f us/file) | read_wav_pcm16 %.3f s (%.0f us/file)"
              % (nf, t_old, t_old / nf * 1e6, t_new, t_new / nf * 1e6))
        S = [np.zeros(((n + 256) // 256, 80), np.float32) for n in lens]
        out = os.path.join(tmp, "o")
        os.makedirs(out)
        t_old = best(lambda: [np.save(os.path.join(out, "a%d" % i), s, allow_pickle=False) for i, s in enumerate(S)])
        t_new = best(lambda: [msf.save_npy(os.path.join(out, "b%d" % i), s) for i, s in enumerate(S)])
        for i in (0, nf // 2, nf - 1):
            assert open(os.path.join(out, "a%d.npy" % i), "rb").read() == open(os.path.join(out, "b%d.npy" % i), "rb").read()
        print("NPY write, %d (T, 80) f32 files, one thread: np.save %.3f s (%.0f us/file) | save_npy %.3f s (%.0f us/file), "
              "bytes identical" % (nf, t_old, t_old / nf * 1e6, t_new, t_new / nf * 1e6))
        for threads in (1, 4, 8):
            st = {}

            def run():
                st.clear()
                msf.make_spect_f0(os.path.join(tmp, "wavs"), os.path.join(tmp, "spmel"), os.path.join(tmp, "raptf0"),
                                  os.path.join(tmp, "spk2gen.pkl"), verbose=False, stats=st, io_threads=threads,
                                  frontend=StubFrontEnd())
            t = best(run)
            print("script flow with a stub front end, io_threads=%d: %.3f s = %.0f files/s around the GPU calls "
                  "(read %.3f, pack %.3f, write %.3f)" % (threads, t, nf / t, st["read_s"], st["pack_s"], st["write_s"]))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
