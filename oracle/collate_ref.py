"""oracle/collate_ref.py - TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

CPU restatement of the reference's training collator, ``data_loader.MyCollator.__call__``
(data_loader.py:101-128), and of the manifest ``make_metadata.py`` writes (:10-33).  PINNED: checked in
tests/test_oracle.py against tests/golden/collate.npz, which the reference's own ``MyCollator`` and
``make_metadata.py`` produced in the build container (tests/golden/make_golden.py; the one line of the
reference that cannot run, ``pdb.set_trace()`` at :106 without an import of pdb, was given a no-op ``pdb``).
"""
import os

import numpy as np


def draw_crop(n_frames, min_len_seq, max_len_seq):
    """data_loader.py:104-105: two draws of two values each from numpy's global generator; only the first
    of each pair is used.  Raises ValueError when the utterance is not longer than the crop (numpy's own)."""
    len_crop = np.random.randint(min_len_seq, max_len_seq + 1, size=2)
    left = np.random.randint(0, n_frames - len_crop[0], size=2)
    return int(left[0]), int(len_crop[0])


def collate(batch, min_len_seq=64, max_len_seq=128, max_len_pad=192):
    """batch: list of (melsp (T,80) f32, emb (82,) f32, f0 (T,) f32) -> numpy arrays
    melsp (B,pad,80) f32 clipped to [0,1] and zero-padded (:111-113), spk_emb (B,82) f32,
    pitch (B,pad,1) f32 padded with -1e10 (:114), len_org (B,) int64 (:116,125)."""
    mels, embs, pitches, lens = [], [], [], []
    for sp, emb, f0 in batch:
        left, n = draw_crop(len(sp), min_len_seq, max_len_seq)
        m = np.zeros((max_len_pad, sp.shape[1]), sp.dtype)
        m[:n] = np.clip(sp[left:left + n, :], 0, 1)
        p = np.full((max_len_pad, 1), -1e10, f0.dtype)
        p[:n, 0] = f0[left:left + n]
        mels.append(m)
        embs.append(emb)
        pitches.append(p)
        lens.append(n)
    return np.stack(mels), np.stack(embs), np.stack(pitches), np.asarray(lens, np.int64)


def metadata(tree):
    """make_metadata.py:10-33 for ``tree`` = {speaker: [file names]}: sorted speakers, each
    [speaker, one-hot (82,) f32 with index 1 for p226 and 7 otherwise, 'speaker/file' in sorted order]."""
    out = []
    for spk in sorted(tree):
        emb = np.zeros((82,), np.float32)
        emb[1 if spk == "p226" else 7] = 1.0
        out.append([spk, emb] + [os.path.join(spk, f) for f in sorted(tree[spk])])
    return out
