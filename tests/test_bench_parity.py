"""bench.py's full-corpus parity sweep (host logic, CPU only): the oracle is run by pool workers over
memory-mapped copies of the GPU outputs; every task re-enters the speaker's MT19937 stream at the right
position.  Here the "GPU outputs" are the oracle's own, so the sweep has to come back exact - and a planted
deviation has to be found."""
import os
import sys

import numpy as np
import torch
from numpy.random import RandomState

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _fake_run():
    from oracle import ref_pipeline as rp
    from speechsplit_b200.corpus import make_manifest, pcm_to_float64, synth_batch
    from speechsplit_b200.sharding import dither_skips, fixed_length
    metas = make_manifest(2, 3, seed=3, mean_s=1.2, std_s=0.2, min_s=1.0, max_s=1.5)
    pcm = synth_batch(metas)
    skips = dither_skips([m.spk for m in metas], [m.length for m in metas])
    off = np.concatenate([[0], np.cumsum([m.length for m in metas])]).astype(np.int64)
    mel, f0n, f0r, bins, fr = [], [], [], [], [0]
    prng, cur = None, None
    for m, p in zip(metas, pcm):
        if m.spk != cur:
            prng, cur = RandomState(m.spk_id), m.spk
        S, fn, st = rp.extract_utterance(pcm_to_float64(p), m.gender, prng, want_stages=True)
        assert len(st["y"]) == fixed_length(m.length)
        mel.append(S), f0n.append(fn), f0r.append(st["f0_rapt"]), bins.append(rp.quantize_f0_numpy(fn)[1])
        fr.append(fr[-1] + len(fn))
    outs = dict(mel=torch.from_numpy(np.concatenate(mel)), f0_norm=torch.from_numpy(np.concatenate(f0n)),
                f0_raw=torch.from_numpy(np.concatenate(f0r)), bins=torch.from_numpy(np.concatenate(bins)))
    return metas, skips, torch.cat(pcm), off, np.asarray(fr, np.int64), outs


def test_parity_sweep_is_exact_on_the_oracles_own_output_and_finds_a_planted_error():
    import bench
    metas, skips, x, off, fr, outs = _fake_run()
    try:
        r = bench.parity_sweep(metas, skips, x, off, fr, outs, "all", per_task=2)   # tasks start mid-stream
        assert r["utterances"] == 6 and r["frames"] == int(fr[-1])
        assert r["mel_max_abs"] == 0.0 and r["identical_bins_frac"] == 1.0 and r["voicing_flag_mismatches"] == 0
        assert r["f0_worst_cents_voiced_both"] == 0.0 and r["utterances_with_any_different_bin"] == 0 and r["pass"]
        outs["mel"][fr[4] + 3, 7] += 3e-4
        voiced = np.nonzero(outs["bins"].numpy()[fr[1]:fr[2]] > 0)[0]
        outs["bins"][fr[1] + voiced[0]] += 1
        r = bench.parity_sweep(metas, skips, x, off, fr, outs, "all", per_task=2)
        assert abs(r["mel_max_abs"] - 3e-4) < 1e-6 and r["mel_values_over_1e-4"] == 1 and not r["pass"]
        assert r["frames_with_different_bin"] == 1 and r["utterances_with_any_different_bin"] == 1
        assert r["worst_utterances"][0]["index"] == 1
        r = bench.parity_sweep(metas, skips, x, off, fr, outs, 2)
        assert r["utterances"] == 2
    finally:
        bench.close_pool()
