"""Host logic of the multi-GPU path, on CPU: LPT sharding, dither stream offsets, and a
world-size-2 gloo run of the same barrier / max-over-ranks pattern bench.py uses."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
from numpy.random import RandomState

from speechsplit_b200.corpus import make_manifest
from speechsplit_b200.sharding import dither_skips, fixed_length, lpt_shards, lpt_shards_by_speaker

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fixed_length():
    assert [fixed_length(n) for n in (255, 256, 257, 48000, 48128)] == [255, 257, 257, 48000, 48129]


def test_dither_skips_follow_the_speaker_stream():
    metas = make_manifest(3, 5, seed=4)
    skips = dither_skips([m.spk for m in metas], [m.length for m in metas])
    for spk in sorted({m.spk for m in metas}):
        idx = [i for i, m in enumerate(metas) if m.spk == spk]
        assert skips[idx[0]] == 0
        prng = RandomState(int(spk[1:]))
        pos = 0
        for i in idx:                       # what make_spect_f0.py:47-55 does, file by file
            assert skips[i] == pos
            n = fixed_length(metas[i].length)
            first = prng.rand(n)[0]
            again = RandomState(int(spk[1:]))
            again.rand(int(skips[i])) if skips[i] else None
            assert again.rand(1)[0] == first
            pos += n


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_lpt_shards_partition_and_balance(world):
    metas = make_manifest(20, 30, seed=1)
    lengths = np.array([m.length for m in metas])
    shards = lpt_shards(lengths, world)
    allidx = np.concatenate(shards)
    assert sorted(allidx.tolist()) == list(range(len(metas)))           # a partition
    loads = np.array([lengths[s].sum() for s in shards])
    assert loads.max() - loads.min() <= lengths.max()                   # LPT bound
    for s in shards:
        assert np.all(np.diff(s) > 0)                                   # corpus order kept inside a shard


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_speaker_atomic_shards(world):
    metas = make_manifest(109, 12, seed=3)
    spk = [m.spk for m in metas]
    lengths = np.array([m.length for m in metas])
    shards = lpt_shards_by_speaker(spk, lengths, world)
    assert sorted(np.concatenate(shards).tolist()) == list(range(len(metas)))
    owner = {}
    for k, s in enumerate(shards):
        for i in s:
            assert owner.setdefault(spk[i], k) == k            # a speaker lives on exactly one shard
        assert np.all(np.diff(s) > 0)
    loads = np.array([lengths[s].sum() for s in shards], dtype=float)
    assert loads.max() / loads.mean() < 1.08
    # fewer speakers than shards: falls back to utterance granularity
    few = make_manifest(2, 10, seed=1)
    sh = lpt_shards_by_speaker([m.spk for m in few], [m.length for m in few], 4)
    assert sorted(np.concatenate(sh).tolist()) == list(range(20)) and all(len(x) for x in sh)


WORKER = r"""
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
from speechsplit_b200.corpus import make_manifest
from speechsplit_b200.sharding import contiguous_shards, dither_skips, lpt_shards
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
metas = make_manifest(6, 10, seed=2)
lengths = [m.length for m in metas]
shard = lpt_shards(lengths, world)[rank]
skips = dither_skips([m.spk for m in metas], lengths)[shard]
mine = torch.tensor([float(sum(lengths[i] for i in shard))], dtype=torch.float64)
tot = mine.clone(); dist.all_reduce(tot)
assert abs(tot.item() - sum(lengths)) < 1e-6                 # every utterance is on exactly one rank
t = torch.tensor([1.0 + rank], dtype=torch.float64)
dist.barrier(); dist.all_reduce(t, op=dist.ReduceOp.MAX)     # bench.py: max over ranks
assert t.item() == float(world)
cnt = torch.tensor([len(shard)]); dist.all_reduce(cnt); assert cnt.item() == len(metas)
# the split bench.py uses: consecutive runs of equal sample count, every rank computes the same cut
cs = contiguous_shards(lengths, world)[rank]
lo = torch.tensor([float(cs[0]) if len(cs) else 0.0, float(len(cs))], dtype=torch.float64)
allr = [torch.zeros(2, dtype=torch.float64) for _ in range(world)]
dist.all_gather(allr, lo)
starts = [int(a[0]) for a in allr]; counts = [int(a[1]) for a in allr]
assert sum(counts) == len(metas) and all(starts[k] + counts[k] == starts[k + 1] for k in range(world - 1))
load = torch.tensor([float(sum(lengths[i] for i in cs))], dtype=torch.float64)
mx = load.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
assert mx.item() <= sum(lengths) / world + max(lengths)
dist.barrier(); dist.destroy_process_group()
print("rank", rank, "ok", len(shard), int(skips.sum()))
"""


def test_two_rank_gloo_sharding(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 2


def test_contiguous_shards_balance_and_stream_order():
    """Consecutive runs of equal sample count: every utterance exactly once, corpus order kept (so each
    shard's requests to a speaker's stream are one ascending stretch), load within one utterance of the
    mean, at most two partial speakers per shard."""
    from speechsplit_b200.corpus import make_manifest
    from speechsplit_b200.sharding import contiguous_shards
    metas = make_manifest(109, 40, seed=0)
    lengths = np.array([m.length for m in metas])
    spk = np.array([m.spk for m in metas])
    for world in (1, 2, 3, 8):
        shards = contiguous_shards(lengths, world)
        assert len(shards) == world
        allidx = np.concatenate(shards)
        assert np.array_equal(allidx, np.arange(len(metas)))
        loads = np.array([lengths[s].sum() for s in shards])
        assert np.abs(loads - lengths.sum() / world).max() <= lengths.max()
        for s in shards:
            partial = [k for k in np.unique(spk[s]) if (spk[s] == k).sum() != (spk == k).sum()]
            assert len(partial) <= 2
    # degenerate: more shards than utterances
    sh = contiguous_shards([5, 7], 4)
    assert sorted(np.concatenate(sh).tolist()) == [0, 1] and len(sh) == 4
