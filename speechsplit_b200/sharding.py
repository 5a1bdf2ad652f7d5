"""Length-balanced sharding of a corpus over the GPUs of one box.

Utterances are independent units: no stage reads another utterance's samples and the F0
statistics are per utterance (make_spect_f0.py:66).  The only coupling in the reference is the
per-speaker dither stream (make_spect_f0.py:47-48,55); it is broken by giving every utterance
its absolute position in that stream (``dither_skip`` = sum of the fixed lengths of the speaker's
earlier files), after which utterances scatter freely.  No collective runs on the hot path.
"""
from typing import List, Sequence

import numpy as np


def fixed_length(n: int) -> int:
    """make_spect_f0.py:52-53."""
    return n + 1 if n % 256 == 0 else n


def dither_skips(speakers: Sequence[str], lengths: Sequence[int]) -> np.ndarray:
    """Stream offset (in doubles) of every utterance, given the corpus in the reference's order
    (speakers sorted, files sorted within a speaker)."""
    skips = np.zeros(len(lengths), dtype=np.uint64)
    pos = {}
    for i, (s, n) in enumerate(zip(speakers, lengths)):
        skips[i] = pos.get(s, 0)
        pos[s] = pos.get(s, 0) + fixed_length(int(n))
    return skips


def lpt_shards(lengths: Sequence[int], n_shards: int) -> List[np.ndarray]:
    """Longest-processing-time-first greedy partition on the sample count (cost is linear in L).
    Each shard keeps corpus order, so requests to a speaker's stream stay sorted."""
    lengths = np.asarray(lengths, dtype=np.int64)
    order = np.argsort(-lengths, kind="stable")
    load = np.zeros(n_shards, dtype=np.int64)
    owner = np.empty(len(lengths), dtype=np.int64)
    for i in order:
        k = int(np.argmin(load))
        owner[i] = k
        load[k] += lengths[i]
    return [np.nonzero(owner == k)[0] for k in range(n_shards)]


def lpt_shards_by_speaker(speakers: Sequence[str], lengths: Sequence[int], n_shards: int) -> List[np.ndarray]:
    """Speaker-atomic variant: whole speakers are dealt to shards, longest first.

    A speaker's MT19937 dither stream is sequential, so a shard that owns only some of a speaker's
    files still has to walk that stream up to its last file.  Keeping speakers whole means every
    stream is generated on exactly one GPU (SURVEY.md 8(e) option 1).  With 109 speakers on 8 GPUs
    the imbalance is at most one speaker (~3 %)."""
    lengths = np.asarray(lengths, dtype=np.int64)
    spk = np.asarray(speakers)
    names, inv = np.unique(spk, return_inverse=True)
    if len(names) < n_shards:           # fewer speakers than shards: fall back to utterance granularity
        return lpt_shards(lengths, n_shards)
    totals = np.bincount(inv, weights=lengths, minlength=len(names))
    owner_of_spk = np.empty(len(names), dtype=np.int64)
    load = np.zeros(n_shards)
    for k in np.argsort(-totals, kind="stable"):
        j = int(np.argmin(load))
        owner_of_spk[k] = j
        load[j] += totals[k]
    owner = owner_of_spk[inv]
    return [np.nonzero(owner == k)[0] for k in range(n_shards)]
