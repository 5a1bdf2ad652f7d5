"""GPU-resident mirror of the reference's ``data_loader.py`` and ``make_metadata.py`` (SURVEY.md 8(f)
ranks 1 and 2): the ``spmel`` / ``raptf0`` NPY trees are read once, kept in HBM, and every training
batch is one ``ssfe_collate`` launch instead of a Python loop of crops, clips and pads.

    reference                                   here
    make_metadata.py:5-33                       make_metadata(root_dir)            (host, unchanged layout)
    data_loader.Utterances       :14-93         Utterances   (same items; features also resident in HBM)
    data_loader.MyCollator       :96-128        MyCollator   (same numpy draws, crop/clip/pad on the GPU)
    data_loader.MultiSampler     :133-152       MultiSampler
    data_loader.get_loader       :157-175       get_loader(hparams) -> torch DataLoader

The batch is the one ``solver.py:142`` unpacks: ``melsp`` f32 (B, max_len_pad, 80) clipped to [0, 1],
``spk_emb`` f32 (B, 82), ``pitch`` f32 (B, max_len_pad, 1) padded with -1e10, ``len_org`` int64 (B,) -
as CUDA tensors, so the solver's ``.to(device)`` calls are no-ops.  The random crop is drawn with the
reference's own two ``np.random.randint`` calls per item in the reference's order (:104-105), so a seeded
run consumes numpy's global stream exactly as the reference does.  (The reference's collator cannot run
as shipped: :106 calls ``pdb.set_trace()`` without importing pdb; that line is the only one not mirrored.)

No CPU fallback: the crop / clip / pad arithmetic exists only as the CUDA kernel.
"""
import os
import pickle

import numpy as np
import torch
from torch.utils import data
from torch.utils.data.sampler import Sampler


def make_metadata(root_dir="assets/spmel", verbose=True):
    """make_metadata.py:5-33: ``<root_dir>/train.pkl`` = one entry per speaker (sorted), each
    ``[speaker, one-hot float32 (82,), 'spk/file.npy', ...]`` with the files in sorted order and the
    reference's hard-coded embedding (index 1 for p226, index 7 for everyone else, :20-24)."""
    dir_name, subdirs, _ = next(os.walk(root_dir))
    if verbose:
        print("Found directory: %s" % dir_name)
    speakers = []
    for speaker in sorted(subdirs):
        if verbose:
            print("Processing speaker: %s" % speaker)
        spkid = np.zeros((82,), dtype=np.float32)
        spkid[1 if speaker == "p226" else 7] = 1.0
        _, _, files = next(os.walk(os.path.join(dir_name, speaker)))
        speakers.append([speaker, spkid] + [os.path.join(speaker, f) for f in sorted(files)])
    with open(os.path.join(root_dir, "train.pkl"), "wb") as handle:
        pickle.dump(speakers, handle)
    return speakers


class _Item(tuple):
    """What ``Utterances.__getitem__`` returns: the reference's ``(melsp, emb_org, f0_org)`` tuple of host
    arrays, plus the item's index so that the collator can address the copy that lives in HBM."""

    def __new__(cls, fields, index):
        self = super().__new__(cls, fields)
        self.index = index
        return self


class Utterances(data.Dataset):
    """data_loader.py:14-93.  As in the reference, an item is a *speaker*: its id, its embedding and the
    features of the first file listed for it (``sbmt[2]``, :62-63); ``mode='train'`` keeps frames
    ``[split:]``, ``'test'`` keeps ``[:split]`` with ``split = 0`` (:21,64-69)."""

    def __init__(self, root_dir, feat_dir, mode, frontend=None):
        self.root_dir, self.feat_dir, self.mode = root_dir, feat_dir, mode
        self.split = 0
        if mode not in ("train", "test"):
            raise ValueError                                          # :48-49
        with open(os.path.join(root_dir, "train.pkl"), "rb") as f:
            meta = pickle.load(f)
        items = []
        for sbmt in meta:
            sp = np.load(os.path.join(root_dir, sbmt[2]))
            f0 = np.load(os.path.join(feat_dir, sbmt[2]))
            if mode == "train":
                sp, f0 = sp[self.split:, :], f0[self.split:]
            else:
                sp, f0 = sp[:self.split, :], f0[:self.split]
            items.append([sbmt[0], sbmt[1], (sp, f0)])
        if mode == "train":
            self.train_dataset = items
        else:
            self.test_dataset = items
        self.num_tokens = len(items)
        self._items = items
        self._fe = frontend
        self._resident = None
        print("Finished loading {} dataset...".format(mode))

    def __getitem__(self, index):
        index = int(index)
        _, emb_org, (melsp, f0_org) = self._items[index]
        return _Item((melsp, emb_org, f0_org), index)

    def __len__(self):
        return self.num_tokens

    @property
    def frontend(self):
        if self._fe is None:
            from .frontend import default_frontend
            self._fe = default_frontend()
        return self._fe

    def resident(self):
        """(mel [sum T, 80] f32, f0 [sum T] f32, frame_offsets int64 [n+1], emb [n, 82] f32): the features
        of all items, uploaded once."""
        if self._resident is None:
            fe = self.frontend
            T = [it[2][0].shape[0] for it in self._items]
            off = np.concatenate([[0], np.cumsum(T)]).astype(np.int64)
            mel = np.concatenate([np.asarray(it[2][0], np.float32).reshape(-1, 80) for it in self._items]) \
                if self._items else np.zeros((0, 80), np.float32)
            f0 = np.concatenate([np.asarray(it[2][1], np.float32).reshape(-1) for it in self._items]) \
                if self._items else np.zeros(0, np.float32)
            emb = np.stack([np.asarray(it[1], np.float32) for it in self._items]) \
                if self._items else np.zeros((0, 82), np.float32)
            self._resident = (fe._dev(mel), fe._dev(f0), off, fe._dev(emb))
        return self._resident


class MyCollator(object):
    """data_loader.py:96-128 with the per-item numpy work replaced by one ``ssfe_collate`` launch."""

    def __init__(self, hparams, dataset=None, want_onehot=False, draws="reference"):
        self.min_len_seq = hparams.min_len_seq
        self.max_len_seq = hparams.max_len_seq
        self.max_len_pad = hparams.max_len_pad
        self.dataset = dataset
        self.want_onehot = want_onehot
        if draws not in ("reference", "batched"):
            raise ValueError("draws must be 'reference' or 'batched'")
        # 'reference': numpy's global generator, two randint calls per item in the reference's order - seeded
        # runs see the reference's batches, at ~8 us per call.  'batched': the same two distributions drawn
        # with two calls per BATCH (a different stream, so not the reference's crops for a given seed).
        self.draws = draws
        self.last_onehot = None          # (onehot (B,pad,257), bins (B,pad)) of the latest batch, if asked for

    def draw(self, batch):
        """The random crops of :104-105, item by item in batch order: (utt, left, len_crop) int arrays."""
        if self.draws == "batched":
            utt = np.asarray([token.index for token in batch], np.int32)
            n_frames = np.asarray([len(token[0]) for token in batch], np.int64)
            len_crop = np.random.randint(self.min_len_seq, self.max_len_seq + 1, size=len(batch)).astype(np.int64)
            if len(batch) and np.any(n_frames - len_crop <= 0):
                raise ValueError("high <= 0")                       # what numpy raises in the reference loop
            left = np.random.randint(0, np.maximum(n_frames - len_crop, 1)) if len(batch) else np.zeros(0, np.int64)
            return utt, np.asarray(left, np.int32), len_crop
        utt, left, len_crop = [], [], []
        for token in batch:
            aa = token[0]
            lc = np.random.randint(self.min_len_seq, self.max_len_seq + 1, size=2)     # 1.5 s ~ 3 s
            lf = np.random.randint(0, len(aa) - lc[0], size=2)
            utt.append(token.index)
            left.append(lf[0])
            len_crop.append(lc[0])
        return np.asarray(utt, np.int32), np.asarray(left, np.int32), np.asarray(len_crop, np.int64)

    def __call__(self, batch):
        if self.dataset is None:
            raise RuntimeError("MyCollator needs the Utterances dataset whose features are resident in HBM")
        utt, left, len_crop = self.draw(batch)
        mel, f0, off, emb = self.dataset.resident()
        fe = self.dataset.frontend
        melsp, pitch, onehot, bins = fe.collate(mel, f0, off, utt, left, len_crop.astype(np.int32),
                                                self.max_len_pad, want_onehot=self.want_onehot)
        self.last_onehot = (onehot, bins) if self.want_onehot else None
        if not len(utt):
            return melsp, emb[:0], pitch, torch.zeros(0, dtype=torch.int64, device=melsp.device)
        # item indices and crop lengths go up as ONE small transfer; from pinned memory and non-blocking, so that
        # the host does not wait for the previous batch's kernels (a pageable copy synchronises the stream)
        meta = torch.empty((2, len(utt)), dtype=torch.int64, pin_memory=emb.is_cuda)
        meta[0] = torch.from_numpy(utt.astype(np.int64))
        meta[1] = torch.from_numpy(len_crop)
        meta = meta.to(emb.device, non_blocking=True)
        return melsp, emb.index_select(0, meta[0]), pitch, meta[1]


class MultiSampler(Sampler):
    """data_loader.py:133-152: every index ``n_repeats`` times per pass, optionally shuffled."""

    def __init__(self, num_samples, n_repeats, shuffle=False):
        self.num_samples = num_samples
        self.n_repeats = n_repeats
        self.shuffle = shuffle

    def gen_sample_array(self):
        idx = torch.arange(self.num_samples, dtype=torch.int64).repeat(self.n_repeats)
        if self.shuffle:
            idx = idx[torch.randperm(len(idx))]
        self.sample_idx_array = idx
        return idx

    def __iter__(self):
        return iter(self.gen_sample_array())

    def __len__(self):
        return self.num_samples * self.n_repeats


def get_loader(hparams, frontend=None, want_onehot=False, draws="reference"):
    """data_loader.py:157-175.  The features live in HBM and a batch is one kernel launch, so the loader
    runs in the main process (``hparams.num_workers`` is the reference's default 0 here whatever it says:
    worker processes could not share the CUDA context) and there is nothing to pin."""
    dataset = Utterances(hparams.root_dir, hparams.feat_dir, hparams.mode, frontend=frontend)
    my_collator = MyCollator(hparams, dataset, want_onehot=want_onehot, draws=draws)
    sampler = MultiSampler(len(dataset), hparams.samplier, shuffle=hparams.shuffle)
    return data.DataLoader(dataset=dataset, batch_size=hparams.batch_size, sampler=sampler, num_workers=0,
                           drop_last=True, pin_memory=False, collate_fn=my_collator)
