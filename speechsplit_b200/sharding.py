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


def contiguous_shards(lengths: Sequence[int], n_shards: int) -> List[np.ndarray]:
    """Cut the corpus, in its own order (speakers sorted, files sorted), into ``n_shards`` consecutive
    runs of equal sample count.

    With the jump-ahead generator (csrc/mt19937.cu) a shard may start anywhere in a speaker's dither
    stream for the price of one segment start, so speakers no longer have to stay whole: every shard holds
    whole speakers plus at most two partial ones, each partial speaker's files are one consecutive stretch
    of its stream, and the load is balanced to within one utterance (the speaker-atomic split of 109
    speakers over 8 GPUs was off by 2.8 %)."""
    lengths = np.asarray(lengths, dtype=np.int64)
    n = len(lengths)
    ends = np.cumsum(lengths)
    total = int(ends[-1]) if n else 0
    cuts = [0]
    for k in range(1, n_shards):
        target = total * k / n_shards
        i = int(np.searchsorted(ends, target))              # first utterance whose end reaches the target
        # cut before or after utterance i, whichever is closer to the target
        before = int(ends[i - 1]) if i > 0 else 0
        cut = i if (target - before) <= (int(ends[min(i, n - 1)]) - target) else i + 1
        cuts.append(min(max(cut, cuts[-1]), n))
    cuts.append(n)
    return [np.arange(cuts[k], cuts[k + 1]) for k in range(n_shards)]
