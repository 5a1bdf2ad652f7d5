#!/usr/bin/env python
"""bench.py - audio-seconds/second of the mel + F0 front end (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Workload (config.workload): BASELINE.json configs[1] - a VCTK-shaped synthetic corpus, 109 speakers x
400 utterances (~3 s avg, 16 kHz, 16-bit PCM).  One step = one pass of the hot path
(make_spect_f0.py:50-74 + utils.quantize_f0_numpy) over the whole corpus.  At N > 1 the SAME corpus
is sharded length-balanced over the ranks (configs[2], strong scaling); there is no collective on
the data path.

  value      audio-s/s, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e        same metric through the C ABI with HOST buffers (pinned H2D of the PCM in, D2H of mel /
             normalised F0 / bins out, inside the timed region)
  roofline   the fused STFT->mel->dB kernel: 1344 algorithmic bytes per frame / its own duration,
             measured with CUDA events on the launch stream inside the timed steps
  cpu_baseline  the reference's CPU arithmetic (scipy filtfilt + numpy RandomState + pySTFT + mel +
             RAPT restatement = oracle/) on a bounded sample, all host cores
  parity     the oracle run over EVERY utterance rank 0 processed (process pool, ~25 s on 16 cores): max and
             99.99th percentile of |d mel|, identical-bin fraction over all frames, utterances with any
             differing bin, voicing-flag mismatches, worst F0 deviation in cents
  configs    (N=1) BASELINE.json configs[0] / [3] / [4] measured in the same process: one 3 s utterance,
             256 x 60 s long form, the batch-16 training front end beside the reference's own torch-eager ops
  --impl reference  times that CPU path alone.
"""
import os

# BLAS / OpenMP threads pinned to 1 for the CPU path (one process per core instead);
# must happen before numpy is imported.  The GPU arm does no host BLAS.
for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
    os.environ.setdefault(_v, "1")

import argparse
import json
import multiprocessing as mp
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FS = 16000
BYTES_PER_FRAME = 1344          # 256 new fp32 samples in + 80 fp32 out
FLOP_PER_FRAME_FFT = 25600      # 2.5 N log2 N, the figure FFT efficiency is quoted against


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--speakers", type=int, default=109)
    ap.add_argument("--utts", type=int, default=400)
    ap.add_argument("--cpu-sample", type=int, default=None, help="utterances in the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--parity-utts", default="all",
                    help="'all' = the oracle over every utterance of rank 0's shard (default), or a number of sampled utterances, 0 = none")
    ap.add_argument("--no-configs", action="store_true", help="skip the configs[0]/[3]/[4] legs of the default run")
    ap.add_argument("--filtfilt-mode", type=int, default=0, help="diagnostics: 1 = sequential validation mode of the filter (one thread per utterance)")
    ap.add_argument("--workload", default="vctk", choices=["vctk", "longform", "single", "collate"],
                    help="vctk = BASELINE configs[1]/[2] (the bench line); longform = configs[3] (256 x 60 s); "
                         "single = configs[0] (one 3 s male utterance); collate = configs[4] (batch-16 training crops)")
    return ap.parse_args()


# ---- CPU path (oracle) --------------------------------------------------------------------------
def _cpu_speaker_job(job):
    """One speaker per task: keeps the per-speaker MT19937 stream semantics (make_spect_f0.py:47)."""
    os.environ["OMP_NUM_THREADS"] = "1"
    from numpy.random import RandomState
    from oracle import ref_pipeline as rp
    spk_id, gender, pcms = job
    prng = RandomState(spk_id)
    secs = 0.0
    for p in pcms:
        x = p.astype(np.float64) / 32768.0
        S, f0n = rp.extract_utterance(x, gender, prng)
        rp.quantize_f0_numpy(f0n)
        secs += len(p) / FS
    return secs


_POOL = None


def pool_size():
    """Rank 0 owns every host core: its CPU legs (baseline, parity sweep) run while the other ranks have
    nothing to do, so the CPU numbers are comparable across N.  The other ranks only build their shard."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    cores = os.cpu_count() or 1
    return max(1, cores if rank == 0 else cores // max(1, world))


def get_pool():
    """Worker processes are forked once, BEFORE this process touches CUDA, and reused."""
    global _POOL
    if _POOL is None:
        _POOL = mp.get_context("fork").Pool(pool_size())
    return _POOL


def cpu_time(jobs):
    pool = get_pool()
    t0 = time.perf_counter()
    secs = sum(pool.map(_cpu_speaker_job, jobs, chunksize=1))
    return secs, time.perf_counter() - t0


def make_cpu_jobs(metas, pcm_of, per_task=25):
    """Group a sample into speaker-atomic tasks of <= per_task utterances."""
    jobs, cur, spk = [], [], None
    for m in metas:
        if spk is not None and (m.spk != spk or len(cur) >= per_task):
            jobs.append((int(spk[1:]), gender, cur))
            cur = []
        spk, gender = m.spk, m.gender
        cur.append(pcm_of(m))
    if cur:
        jobs.append((int(spk[1:]), gender, cur))
    return jobs


# ---- clocks -------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock, power and throttle reasons DURING the timed region.  NVML is polled from a thread every
    few milliseconds (a timed region at N=8 lasts ~65 ms, too short for `nvidia-smi -lms`); nvidia-smi is
    the fallback when the NVML binding is missing."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NVML_REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index):
        self.p = self.f = self.thread = None
        self.samples = []
        try:
            import threading

            import pynvml
            import torch
            pynvml.nvmlInit()
            try:
                uuid = str(torch.cuda.get_device_properties(gpu_index).uuid)
                h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:
                h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.nv, self.h = pynvml, h
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self.stop_flag = threading.Event()

            def poll():
                while not self.stop_flag.is_set():
                    try:
                        sm = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                        pw = pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0
                        try:
                            rs = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                        except Exception:
                            rs = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                        self.samples.append((sm, pw, int(rs)))
                    except Exception:
                        pass
                    self.stop_flag.wait(0.004)

            self.thread = threading.Thread(target=poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.thread is not None:
            self.stop_flag.set()
            self.thread.join(timeout=2)
            if self.samples:
                sm = sorted(x[0] for x in self.samples)
                reasons = set()
                for _, _, rs in self.samples:
                    for bit, nme in self.NVML_REASONS.items():
                        if rs & bit:
                            reasons.add(nme)
                out.update(sm_mhz=float(np.median(sm[len(sm) // 2:])), sm_max_mhz=self.mx, reasons=sorted(reasons),
                           samples=len(sm), power_w_max=float(max(x[1] for x in self.samples)), source="nvml")
            return out
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])), mx.append(float(c[2])), pw.append(float(c[3]))
            except ValueError:
                continue
            for nme, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        self.f.close()
        os.unlink(self.f.name)
        if sm:
            hi = sorted(sm)[len(sm) // 2:]          # upper half = samples under load
            out.update(sm_mhz=float(np.median(hi)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons),
                       samples=len(sm), power_w_max=float(max(pw)), source="nvidia-smi")
        return out


# ---- corpus ---------------------------------------------------------------------------------------
def build_shard(args, rank, world):
    """Manifest -> LPT shard of this rank -> control tracks (process pool, before CUDA is touched)."""
    from speechsplit_b200.corpus import control_tracks, make_manifest
    from speechsplit_b200.sharding import contiguous_shards, dither_skips
    if args.workload == "longform":       # configs[3]: 4 speakers x 64 utterances of 60.000 s
        metas = make_manifest(4, 64, seed=0, fixed_len=960000)
    elif args.workload == "single":       # configs[0]: one 3 s male utterance (p226)
        metas = make_manifest(1, 1, first_id=226, seed=0, fixed_len=48000)
    else:
        metas = make_manifest(args.speakers, args.utts, seed=0)
    skips = dither_skips([m.spk for m in metas], [m.length for m in metas])
    shard = contiguous_shards([m.length for m in metas], world)[rank]
    mine = [metas[i] for i in shard]
    tracks = get_pool().map(control_tracks, mine, chunksize=64)
    return metas, mine, skips[shard], tracks


def synth_on_gpu(mine, tracks, device, chunk=768):
    """Synthesise the shard on the GPU (length-sorted batches), return one int16 tensor + offsets."""
    import torch
    from speechsplit_b200.corpus import synth_batch
    order = np.argsort([m.length for m in mine], kind="stable")
    lengths = np.array([m.length for m in mine], dtype=np.int64)
    off = np.concatenate([[0], np.cumsum(lengths)]).astype(np.int64)
    x = torch.empty(int(off[-1]), dtype=torch.int16, device=device)
    for s in range(0, len(order), chunk):
        idx = order[s:s + chunk]
        pcm = synth_batch([mine[i] for i in idx], device=device, tracks=[tracks[i] for i in idx])
        for i, p in zip(idx, pcm):
            x[off[i]:off[i + 1]] = p
        del pcm
    torch.cuda.synchronize()
    return x, off


# ---- the other BASELINE.json configurations, measured in the same process (N=1) ------------------------
def _dev_timer(fn, n, torch):
    """(device ms per call, wall ms per call) of n back-to-back asynchronous calls."""
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (time.perf_counter() - t0) * 1e3 / n


def leg_single(fe, dev):
    """configs[0]: make_spect_f0 on one 3 s male utterance (p226) - the latency of one call."""
    import torch
    from numpy.random import RandomState
    from oracle import ref_pipeline as rp
    from speechsplit_b200.corpus import UttMeta, pcm_to_float64, synth_batch
    meta = UttMeta("p226", "M", 0, 48000, 226000)
    pcm = synth_batch([meta], device=dev)[0]
    off, lo, hi, seed, skip = [0, 48000], [50.0], [250.0], [226], [0]
    want = ("mel", "f0_norm", "f0_raw", "bins", "onehot")
    res = fe.extract(pcm, off, lo, hi, seed, skip, want=want)
    out = {k: res[k] for k in want}

    def call():
        fe.extract(pcm, off, lo, hi, seed, skip, want=want, out=out)

    for _ in range(20):
        call()
    l0 = fe.launch_count
    dev_ms, wall_ms = _dev_timer(call, 200, torch)
    launches = (fe.launch_count - l0) / 200
    # one synchronous call, as a utils.* drop-in user makes it (results read back)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(50):
        call()
        out["bins"][0].item()
    sync_ms = (time.perf_counter() - t0) * 1e3 / 50
    x = pcm_to_float64(pcm)
    t0 = time.perf_counter()
    for _ in range(3):
        S, f0n, st = rp.extract_utterance(x, "M", RandomState(226), want_stages=True)
        rb = rp.quantize_f0_numpy(f0n)[1]
    cpu_ms = (time.perf_counter() - t0) * 1e3 / 3
    mel = out["mel"].cpu().numpy()
    same = float((out["bins"].cpu().numpy() == rb).mean())
    return {"workload": "BASELINE configs[0]: one 3 s male utterance (p226), lo/hi 50/250", "ms_per_call_device": dev_ms,
            "ms_per_call_host_async": wall_ms, "ms_per_call_sync_with_readback": sync_ms, "launches_per_call": launches,
            "audio_s_per_s": 3.0 / (dev_ms * 1e-3), "cpu_ms_per_call_1core": cpu_ms,
            "parity": {"mel_max_abs": float(np.abs(mel - S).max()), "identical_bins_frac": same,
                       "pass": bool(np.abs(mel - S).max() <= 1e-4 and same >= 0.999)}}


def leg_longform(fe, dev, steps=3):
    """configs[3]: 256 utterances of 60.000 s (4 speakers x 64): stresses the filter scan and the Viterbi chain."""
    import torch
    from speechsplit_b200.corpus import control_tracks, make_manifest
    from speechsplit_b200.sharding import dither_skips
    metas = make_manifest(4, 64, seed=0, fixed_len=960000)
    skips = dither_skips([m.spk for m in metas], [m.length for m in metas])
    tracks = get_pool().map(control_tracks, metas, chunksize=4)
    x, off = synth_on_gpu(metas, tracks, dev, chunk=32)
    del tracks
    lo = np.array([50.0 if m.gender == "M" else 100.0 for m in metas], np.float32)
    hi = np.array([250.0 if m.gender == "M" else 600.0 for m in metas], np.float32)
    seed = np.array([m.spk_id for m in metas], np.uint32)
    fix, fr = fe.plan(off)
    T = int(fr[-1])
    want = ("mel", "f0_norm", "f0_raw", "bins", "onehot")
    outs = dict(mel=torch.empty((T, 80), dtype=torch.float32, device=dev), f0_norm=torch.empty(T, dtype=torch.float32, device=dev),
                f0_raw=torch.empty(T, dtype=torch.float32, device=dev), bins=torch.empty(T, dtype=torch.int64, device=dev),
                onehot=torch.empty((T, 257), dtype=torch.float32, device=dev))

    def step():
        fe.extract(x, off, lo, hi, seed, skips, want=want, out=outs)

    for _ in range(3):
        step()
    fe.enable_timing(True)
    dev_ms, _ = _dev_timer(step, steps, torch)
    stages = fe.stage_ms()
    fe.enable_timing(False)
    audio_s = float(off[-1]) / FS
    xh = torch.empty(x.shape, dtype=torch.int16, pin_memory=True)
    xh.copy_(x)
    ho = dict(mel=torch.empty((T, 80), dtype=torch.float32, pin_memory=True), f0_norm=torch.empty(T, dtype=torch.float32, pin_memory=True),
              bins=torch.empty(T, dtype=torch.int64, pin_memory=True))
    for _ in range(2):
        fe.extract_host(xh, off, lo, hi, seed, skips, out=ho)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        fe.extract_host(xh, off, lo, hi, seed, skips, out=ho)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / steps
    del xh, ho
    parity = parity_sweep(metas, skips, x, off, fr, outs, 16, per_task=1)
    return {"workload": "BASELINE configs[3]: 256 x 60.000 s utterances (4 speakers x 64), %.0f audio-s, %d frames" % (audio_s, T),
            "ms_per_step": dev_ms, "audio_s_per_s": audio_s / (dev_ms * 1e-3), "e2e_ms_per_step": e2e_ms,
            "e2e_audio_s_per_s": audio_s / (e2e_ms * 1e-3), "stage_ms": stages, "parity": parity}


def leg_collate(fe, dev, mine, fr, outs, cpu=True, n_steps=200):
    """configs[4]: the online training front end of solver.py:142-163 at batch 16 - crop (64..128 frames) + clip +
    pad to 192 (data_loader.py:101-128), InterpLnr (model.py:380-436), quantize_f0_torch (utils.py:62-74) - with
    the features resident in HBM.  Opponent: the reference's own torch-eager ops on the SAME GPU
    (oracle/torch_eager_ref.py), once starting from the collated batch on the device and once including the
    reference's CPU collator + H2D."""
    import shutil
    from types import SimpleNamespace

    import torch
    from oracle import collate_ref
    from oracle.torch_eager_ref import interp_lnr_eager, quantize_f0_eager
    from speechsplit_b200 import utils as su
    from speechsplit_b200.data_loader import get_loader, make_metadata
    from speechsplit_b200.interp import InterpLnr
    rng = np.random.default_rng(0)
    frs = np.diff(fr)
    cand = np.nonzero(frs > 130)[0]
    interp = InterpLnr().to(dev).train()

    def draw():
        utt = rng.choice(cand, 16)
        ln = rng.integers(64, 129, 16)
        left = np.array([rng.integers(0, frs[u] - l) for u, l in zip(utt, ln)])
        return utt, left, ln

    def ours():
        utt, left, ln = draw()
        melsp, pitch, _, _ = fe.collate(outs["mel"], outs["f0_norm"], fr, utt, left, ln, 192, want_onehot=False)
        xi = interp(torch.cat((melsp, pitch), dim=-1), torch.from_numpy(ln).to(dev))      # solver.py:160-161
        return su.quantize_f0_torch(xi[:, :, -1])[0]                                    # solver.py:162

    utt0, left0, ln0 = draw()
    melsp0, pitch0, _, _ = fe.collate(outs["mel"], outs["f0_norm"], fr, utt0, left0, ln0, 192, want_onehot=False)
    len0 = torch.from_numpy(ln0).to(dev)

    def eager_ops():                      # the reference's GPU work for one batch that is already on the device
        xi = interp_lnr_eager(torch.cat((melsp0, pitch0), dim=-1), len0)
        return quantize_f0_eager(xi[:, :, -1])[0]

    mel_h, f0_h = outs["mel"].cpu().numpy(), outs["f0_norm"].cpu().numpy()
    items = [(mel_h[fr[u]:fr[u + 1]], np.zeros(82, np.float32), f0_h[fr[u]:fr[u + 1]]) for u in cand[:64]]

    def eager_full():                     # + the reference's CPU collator (num_workers=0) and the H2D of its batch
        bt = [items[j] for j in rng.integers(0, len(items), 16)]
        m, e, pch, ln = collate_ref.collate(bt, 64, 128, 192)
        m, pch, ln = torch.from_numpy(m).to(dev), torch.from_numpy(pch).to(dev), torch.from_numpy(ln).to(dev)
        xi = interp_lnr_eager(torch.cat((m, pch), dim=-1), ln)
        return quantize_f0_eager(xi[:, :, -1])[0]

    res = {"batch": 16, "max_len_pad": 192,
           "step": "crop + clip + pad (data_loader.py:101-128) + InterpLnr (model.py:380-436) + quantize_f0_torch (utils.py:62-74), solver.py:142-163"}
    for name, fn in (("ours", ours), ("reference_eager_ops_same_gpu", eager_ops), ("reference_collator_cpu_plus_eager_same_gpu", eager_full)):
        for _ in range(20):
            fn()
        # these steps are bound by the host (Python, numpy draws, launches): the best of three rounds, for every leg
        # alike, keeps another tenant's burst on the box's cores out of the ratio
        d_ms, w_ms = min((_dev_timer(fn, n_steps, torch) for _ in range(3)), key=lambda t: max(t))
        res[name + "_steps_per_s"] = 1e3 / max(d_ms, w_ms)
        res[name + "_ms_device"] = d_ms
        res[name + "_ms_wall"] = w_ms
    res["speedup_vs_reference_eager_ops"] = res["ours_steps_per_s"] / res["reference_eager_ops_same_gpu_steps_per_s"]
    res["speedup_vs_reference_collator_plus_eager"] = res["ours_steps_per_s"] / res["reference_collator_cpu_plus_eager_same_gpu_steps_per_s"]
    # same draws -> same result as the reference's eager ops (bit-exact; the unit tests hold the golden vectors)
    torch.manual_seed(0)
    sc, lsg = interp.draw(16, dev)
    a = interp.resample(torch.cat((melsp0, pitch0), dim=-1), len0, sc, lsg)
    b = interp_lnr_eager(torch.cat((melsp0, pitch0), dim=-1), len0, draws=(sc, lsg))
    qa, qb = su.quantize_f0_torch(a[:, :, -1])[0], quantize_f0_eager(b[:, :, -1])[0]
    res["parity"] = {"interp_bit_exact": bool(torch.equal(a, b)), "onehot_bit_exact": bool(torch.equal(qa, qb))}

    # the same step through the reference-facing loader API (speechsplit_b200.data_loader.get_loader over
    # spmel / raptf0 NPY trees + train.pkl): one item per speaker = its first file, as data_loader.py:62-63
    tmp = tempfile.mkdtemp(prefix="ssfe_loader_")
    try:
        seen = set()
        for i, m in enumerate(mine):
            if m.spk in seen or frs[i] <= 130:
                continue
            seen.add(m.spk)
            for sub, t in (("spmel", mel_h), ("raptf0", f0_h)):
                os.makedirs(os.path.join(tmp, sub, m.spk), exist_ok=True)
                np.save(os.path.join(tmp, sub, m.spk, "%s_001.npy" % m.spk), t[fr[i]:fr[i + 1]], allow_pickle=False)
        make_metadata(os.path.join(tmp, "spmel"), verbose=False)
        n_batches = 220
        hp = SimpleNamespace(root_dir=os.path.join(tmp, "spmel"), feat_dir=os.path.join(tmp, "raptf0"), mode="train",
                             batch_size=16, shuffle=True, num_workers=0, samplier=-(-16 * n_batches // len(seen)),
                             min_len_seq=64, max_len_seq=128, max_len_pad=192)
        loader = get_loader(hp, frontend=fe, want_onehot=False)
        it = iter(loader)

        def loader_step():
            melsp, emb, pitch, len_org = next(it)                              # solver.py:142
            xi = interp(torch.cat((melsp, pitch), dim=-1), len_org)           # solver.py:160-161
            return su.quantize_f0_torch(xi[:, :, -1])[0]

        for _ in range(20):
            loader_step()
        d_ms, w_ms = _dev_timer(loader_step, n_batches - 20 - 1, torch)
        res["loader_steps_per_s"] = 1e3 / max(d_ms, w_ms)
        res["loader"] = ("speechsplit_b200.data_loader.get_loader, %d speakers, features resident in HBM, draws in the "
                         "reference's order (two np.random.randint calls per item)" % len(seen))
        if cpu:
            # the reference's collator loop + quantisation on one host core (its loader default is num_workers=0)
            from oracle import ref_pipeline
            t0, nb = time.perf_counter(), 0
            while time.perf_counter() - t0 < 2.0:
                bt = [items[j] for j in rng.integers(0, len(items), 16)]
                _, _, pitch_ref, _ = collate_ref.collate(bt, 64, 128, 192)
                ref_pipeline.quantize_f0_numpy(pitch_ref.reshape(-1))
                nb += 1
            res["cpu_collator_steps_per_s_1core"] = nb / (time.perf_counter() - t0)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return res


def leg_script(fe, dev, n_speakers=109, utts=40):
    """the script form end to end - a WAV tree (109 speakers x 40 files) + spk2gen.pkl in,
    the spmel / raptf0 NPY trees out (make_spect_f0.py:19-74) - and where its wall time goes."""
    import pickle
    import shutil
    import wave

    from speechsplit_b200 import make_spect_f0 as msf
    from speechsplit_b200.corpus import control_tracks, make_manifest
    metas = make_manifest(n_speakers, utts, seed=7)
    tracks = get_pool().map(control_tracks, metas, chunksize=64)
    x, off = synth_on_gpu(metas, tracks, dev)
    pcm = x.cpu().numpy()
    base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
    tmp = tempfile.mkdtemp(prefix="ssfe_script_", dir=base)
    try:
        spk2gen = {}
        for i, m in enumerate(metas):
            d = os.path.join(tmp, "wavs", m.spk)
            os.makedirs(d, exist_ok=True)
            spk2gen[m.spk] = m.gender
            with wave.open(os.path.join(d, "%s_%03d.wav" % (m.spk, m.index + 1)), "wb") as w:
                w.setnchannels(1)
                w.setsampwidth(2)
                w.setframerate(FS)
                w.writeframes(pcm[off[i]:off[i + 1]].tobytes())
        with open(os.path.join(tmp, "spk2gen.pkl"), "wb") as fh:
            pickle.dump(spk2gen, fh)

        def run(tag):
            st = {}
            t0 = time.perf_counter()
            msf.make_spect_f0(os.path.join(tmp, "wavs"), os.path.join(tmp, "spmel" + tag), os.path.join(tmp, "raptf0" + tag),
                              os.path.join(tmp, "spk2gen.pkl"), device=dev.index, verbose=False, stats=st)
            st["wall_s"] = time.perf_counter() - t0
            return st

        run("_warm")
        st = run("")
        n_files = int(st.get("files", 0))
        audio_s = float(off[-1]) / FS
        # one file back through numpy, as data_loader.py:62-63 reads it
        m0 = metas[0]
        a = np.load(os.path.join(tmp, "spmel", m0.spk, "%s_%03d.npy" % (m0.spk, 1)))
        ok = bool(a.dtype == np.float32 and a.ndim == 2 and a.shape[1] == 80)
        return {"workload": "script form: %d speakers x %d WAV files (16-bit PCM, %.0f audio-s) -> spmel/ + raptf0/ NPY trees, files on %s"
                            % (n_speakers, utts, audio_s, "tmpfs" if base else "disk"),
                "files": n_files, "files_per_s": n_files / st["wall_s"], "audio_s_per_s": audio_s / st["wall_s"],
                "wall_s": st["wall_s"], "read_wav_s": st.get("read_s"), "pack_s": st.get("pack_s"),
                "gpu_extract_host_s": st.get("extract_s"), "write_npy_s": st.get("write_s"), "extract_calls": st.get("calls"),
                "other_python_s": st["wall_s"] - sum(st.get(k, 0.0) for k in ("read_s", "pack_s", "extract_s", "write_s")),
                "npy_readback_ok": ok}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


# ---- our arm --------------------------------------------------------------------------------------
def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    if args.gpus > 1 and world == 1:
        raise SystemExit("launch multi-GPU runs with torch.distributed.run (one rank per GPU)")
    metas, mine, skips, tracks = build_shard(args, rank, world)

    import torch
    import torch.distributed as dist
    from speechsplit_b200 import FrontEnd, FrontEndConfig
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    x, off = synth_on_gpu(mine, tracks, dev)
    del tracks
    n = len(mine)
    lo = np.array([50.0 if m.gender == "M" else 100.0 for m in mine], np.float32)
    hi = np.array([250.0 if m.gender == "M" else 600.0 for m in mine], np.float32)
    seed = np.array([m.spk_id for m in mine], np.uint32)
    audio_s_rank = float((off[-1]) / FS)
    audio_s_total = float(sum(m.length for m in metas) / FS)

    fe = FrontEnd(local, FrontEndConfig(filtfilt_mode=args.filtfilt_mode))
    fix, fr = fe.plan(off)
    T = int(fr[-1])
    outs = dict(mel=torch.empty((T, 80), dtype=torch.float32, device=dev),
                f0_norm=torch.empty(T, dtype=torch.float32, device=dev),
                f0_raw=torch.empty(T, dtype=torch.float32, device=dev),
                bins=torch.empty(T, dtype=torch.int64, device=dev),
                onehot=torch.empty((T, 257), dtype=torch.float32, device=dev))
    want = ("mel", "f0_norm", "f0_raw", "bins", "onehot")

    def step():
        return fe.extract(x, off, lo, hi, seed, skips, want=want, out=outs)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    # ---- timed region: K steps, device timed, stage events on the launch stream ------------------
    fe.enable_timing(True)
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = fe.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage_acc = {}
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step()                                      # asynchronous: no host sync inside the timed region
    ev1.record()
    barrier()
    stage_acc = {k: v * args.steps for k, v in fe.stage_ms().items()}   # averaged over the timed calls
    ms = ev0.elapsed_time(ev1)
    launches = fe.launch_count - launches0
    clocks = sampler.stop() if sampler else None
    fe.enable_timing(False)
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(t_ms.item()) / args.steps
    value = audio_s_total / (ms_per_step * 1e-3)
    t_fr = torch.tensor([T], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_fr)
    T_total = int(t_fr.item())

    # ---- end to end through the C ABI with host buffers --------------------------------------------
    e2e = None
    if not args.no_e2e:
        xh = torch.empty(x.shape, dtype=torch.int16, pin_memory=True)
        xh.copy_(x)
        ho = dict(mel=torch.empty((T, 80), dtype=torch.float32, pin_memory=True),
                  f0_norm=torch.empty(T, dtype=torch.float32, pin_memory=True),
                  bins=torch.empty(T, dtype=torch.int64, pin_memory=True))
        for _ in range(2):
            fe.extract_host(xh, off, lo, hi, seed, skips, out=ho)
        barrier()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            fe.extract_host(xh, off, lo, hi, seed, skips, out=ho)     # returns with results on the host
        e1.record()
        barrier()
        wall = (time.perf_counter() - t0) * 1e3
        tm = torch.tensor([max(e0.elapsed_time(e1), wall)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        e2e_ms = float(tm.item()) / args.steps
        h2d = int(x.numel() * 2)
        d2h = int(T * 80 * 4 + T * 4 + T * 8)
        tb = torch.tensor([h2d, d2h], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tb)
        e2e = {"value": audio_s_total / (e2e_ms * 1e-3), "unit": "audio-s/s", "ms_per_step": e2e_ms,
               "h2d_bytes_per_step": int(tb[0].item()), "d2h_bytes_per_step": int(tb[1].item()),
               "host_input": "int16 PCM (pinned)", "host_output": "mel f32 + f0_norm f32 + bins i64 (pinned)"}
        same = bool(np.array_equal(ho["mel"][:4096].numpy(), outs["mel"][:4096].cpu().numpy()))
        e2e["matches_device_path"] = same
        # the ceiling of this box for exactly these transfers: the step's H2D and D2H bytes, both directions at
        # once, ALL ranks at once, nothing else running (max over ranks, like the step itself)
        s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()

        def both_ways():
            with torch.cuda.stream(s_up):
                x.copy_(xh, non_blocking=True)
            with torch.cuda.stream(s_dn):
                ho["mel"].copy_(outs["mel"], non_blocking=True)
                ho["f0_norm"].copy_(outs["f0_norm"], non_blocking=True)
                ho["bins"].copy_(outs["bins"], non_blocking=True)

        both_ways()
        barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            both_ways()
        torch.cuda.synchronize()
        tr = torch.tensor([(time.perf_counter() - t0) / 3 * 1e3], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tr, op=dist.ReduceOp.MAX)
        roof_ms = float(tr.item())
        e2e["host_roof_ms"] = roof_ms
        e2e["host_roof_gbs"] = (e2e["h2d_bytes_per_step"] + e2e["d2h_bytes_per_step"]) / (roof_ms * 1e-3) / 1e9
        e2e["host_roof_value"] = audio_s_total / (roof_ms * 1e-3)
        e2e["frac_of_host_roof"] = e2e["value"] / e2e["host_roof_value"]
        e2e["host_roof_note"] = ("pinned host <-> device copies of this step's bytes alone, both directions concurrently on all %d "
                                 "rank(s); e2e can at best equal it (every byte crosses the link once)" % world)
        del xh, ho

    collate = None
    if args.workload == "collate" and rank == 0:
        collate = leg_collate(fe, dev, mine, fr, outs, cpu=not args.no_cpu_baseline)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the fused kernel (rank 0's launches) -------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    stft_ms = stage_acc["stft_mel"] / args.steps
    achieved = T * BYTES_PER_FRAME / (stft_ms * 1e-3) / 1e9
    # `achieved` / `peak` / `frac` are the HBM figures the contract asks for (algorithmic bytes over the measured
    # copy peak); `bound` names what actually limits the kernel (ncu: shared-memory wavefronts and FP32 issue, DRAM
    # ~10 % busy), and the FP32 fraction is carried beside it.
    roofline = {"kernel": "stft_mel_kernel<0> (fused STFT->mel->dB->normalise)", "bound": "smem/fp32-issue",
                "denominator": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s",
                "traffic": None, "frames_per_launch": T, "bytes_per_frame": BYTES_PER_FRAME,
                "launch_ms": stft_ms, "fft_tflops_nominal": T * FLOP_PER_FRAME_FFT / (stft_ms * 1e-3) / 1e12}
    # 70.6 TFLOP/s = the FMA rate measured with profiles/microbench/fp64_bench.cu on this pool's B200s
    roofline["fp32_peak_tflops"] = 70.6
    roofline["fft_frac_of_fp32_peak"] = roofline["fft_tflops_nominal"] / 70.6
    tr = os.path.join(ROOT, "profiles", "stft_traffic.json")
    if os.path.exists(tr):
        try:
            t = json.load(open(tr))
            # measured by ncu (dram__bytes_read.sum + dram__bytes_write.sum of one launch) at this bench configuration;
            # scaled by the frame count only if the capture had a different one
            if int(t.get("frames", 0)) == T:
                roofline["traffic"] = t["dram_bytes_per_launch"]
            else:
                roofline["traffic"] = t["dram_bytes_per_frame"] * T
            roofline["traffic_source"] = t.get("source")
            roofline["note"] = t.get("note")
        except Exception:
            pass
    stages = {k: v / args.steps for k, v in stage_acc.items()}

    # ---- parity gate on a sample of this very corpus -------------------------------------------------
    parity = parity_sweep(mine, skips, x, off, fr, outs, args.parity_utts if args.parity_utts == "all" else int(args.parity_utts))

    # ---- CPU baseline on a bounded sample ---------------------------------------------------------------
    cpu = None
    if not args.no_cpu_baseline:
        cpu = cpu_baseline(mine, x, off, args)

    # ---- BASELINE configs[0] / [3] / [4] in the same run (N=1 only: they are single-GPU configurations) -------
    configs = None
    if world == 1 and args.workload == "vctk" and not args.no_configs:
        configs = {}
        for name, fn in (("single", lambda: leg_single(fe, dev)),
                         ("collate", lambda: leg_collate(fe, dev, mine, fr, outs, cpu=not args.no_cpu_baseline)),
                         ("longform", lambda: leg_longform(fe, dev)),
                         ("script", lambda: leg_script(fe, dev))):
            try:
                configs[name] = fn()
            except Exception as ex:           # a failed side leg must not take the bench line with it
                configs[name] = {"error": "%s: %s" % (type(ex).__name__, ex)}

    line = {"metric": "audio-sec/sec mel+F0", "value": value, "unit": "audio-s/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32 (STFT/mel), f64 (filtfilt), f32+f64 (RAPT)",
            "data": "synthetic",
            "config": {"workload": "VCTK-shaped synthetic corpus: %d speakers x %d utterances, %.0f audio-s, 16 kHz int16 PCM "
                                   "(BASELINE.json configs[1]); at N>1 the same corpus cut into N consecutive runs of equal sample count (configs[2])"
                                   % (args.speakers, args.utts, audio_s_total),
                       "utterances": len(metas), "frames": int(T_total),
                       "outputs": "mel f32 [T,80], f0_norm f32 [T], bins i64 [T], one-hot f32 [T,257]",
                       "l2": "inputs per step (%.1f GB PCM on rank 0) exceed the 126 MB L2; no flush needed" % (x.numel() * 2 / 1e9),
                       "collective": "none on the data path"},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
            "cpu_baseline": cpu, "stage_ms": stages, "parity": parity, "configs": configs}
    if args.workload != "vctk":
        line["config"]["workload"] = {"longform": "BASELINE configs[3]: 256 x 60.000 s utterances (4 speakers x 64)",
                                      "single": "BASELINE configs[0]: one 3 s male utterance (p226)",
                                      "collate": line["config"]["workload"] + " + configs[4] collator"}[args.workload]
    if collate:
        line["collate"] = collate
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# |d mel| histogram of the parity sweep: 10 bins per decade from 1e-8 to 1e-2 (+ underflow / overflow)
_HIST_EDGES = np.concatenate([[0.0], 10.0 ** np.arange(-8.0, -1.95, 0.1), [np.inf]])


def _parity_job(job):
    """Oracle vs GPU for a run of consecutive utterances of ONE speaker (the speaker's MT19937 stream is
    advanced to the first of them, then continues as in make_spect_f0.py:47-55).  GPU results and PCM are read
    from memory-mapped .npy files (the pool was forked before CUDA existed)."""
    os.environ["OMP_NUM_THREADS"] = "1"
    from numpy.random import RandomState
    from oracle import ref_pipeline as rp
    shm, spk_id, gender, utts = job
    pcm = np.load(os.path.join(shm, "pcm.npy"), mmap_mode="r")
    mel = np.load(os.path.join(shm, "mel.npy"), mmap_mode="r")
    bins = np.load(os.path.join(shm, "bins.npy"), mmap_mode="r")
    f0n_g = np.load(os.path.join(shm, "f0_norm.npy"), mmap_mode="r")
    f0r_g = np.load(os.path.join(shm, "f0_raw.npy"), mmap_mode="r")
    prng = RandomState(spk_id)
    pos = 0
    unv = np.float32(-1e10)
    r = dict(frames=0, same_bins=0, utts=0, utts_with_diff=0, voicing_mismatch=0, mel_max=0.0, mel_over=0,
             cents_max=0.0, voiced_both=0, cents_over=0, f0_norm_max=0.0, audio_s=0.0,
             hist=np.zeros(len(_HIST_EDGES) - 1, np.int64), worst=[], band_max=np.zeros(80), cells=[])
    for (idx, s0, s1, f0, f1, skip) in utts:
        while pos < skip:                      # advance the stream (bounded chunks)
            n = int(min(skip - pos, 1 << 22))
            prng.rand(n)
            pos += n
        x = np.asarray(pcm[s0:s1]).astype(np.float64) / 32768.0
        S, f0n, st = rp.extract_utterance(x, gender, prng, want_stages=True)
        pos += len(st["y"])
        gm = np.asarray(mel[f0:f1])
        assert gm.shape == S.shape, (gm.shape, S.shape)
        d = np.abs(gm - S)
        r["hist"] += np.histogram(d, bins=_HIST_EDGES)[0]
        dm = float(d.max())
        r["mel_max"] = max(r["mel_max"], dm)
        r["band_max"] = np.maximum(r["band_max"], d.max(axis=0))
        if dm > 5e-5:
            t, bnd = np.unravel_index(int(d.argmax()), d.shape)
            r["cells"].append((dm, int(idx), int(t), int(bnd), float(S[t, bnd]), float(gm[t, bnd]), int(S.shape[0])))
        r["mel_over"] += int((d > 1e-4).sum())
        rb = rp.quantize_f0_numpy(f0n)[1]
        gb = np.asarray(bins[f0:f1])
        same = int((gb == rb).sum())
        r["frames"] += rb.size
        r["same_bins"] += same
        r["utts"] += 1
        r["utts_with_diff"] += int(same != rb.size)
        gr = np.asarray(f0r_g[f0:f1])
        rr = st["f0_rapt"]
        vg, vr = gr != unv, rr != unv
        r["voicing_mismatch"] += int((vg != vr).sum())
        both = vg & vr
        if both.any():
            c = 1731.234 * np.abs(gr[both].astype(np.float64) - rr[both].astype(np.float64))   # cents = 1200 log2(f / f_ref)
            r["cents_max"] = max(r["cents_max"], float(c.max()))
            r["cents_over"] += int((c > 1.0).sum())
            r["voiced_both"] += int(both.sum())
        gn = np.asarray(f0n_g[f0:f1])
        bn = (gn > 0) & (f0n > 0)
        if bn.any():
            r["f0_norm_max"] = max(r["f0_norm_max"], float(np.abs(gn[bn] - f0n[bn]).max()))
        r["audio_s"] += (s1 - s0) / FS
        if same != rb.size or dm > 5e-5:
            r["worst"].append((int(idx), rb.size - same, dm))
    return r


def parity_sweep(mine, skips, x, off, fr, outs, which, per_task=40):
    """The parity gates of SURVEY.md 8(d) over `which` = 'all' utterances of this rank's shard (or n sampled)."""
    import shutil
    if which in (0, "0", None):
        return None
    n = len(mine)
    if which == "all":
        idx = np.arange(n)
    else:
        idx = np.unique(np.linspace(0, n - 1, int(which)).astype(int))
    base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
    shm = tempfile.mkdtemp(prefix="ssfe_parity_", dir=base)
    t0 = time.perf_counter()
    try:
        np.save(os.path.join(shm, "pcm.npy"), x.cpu().numpy())
        for k in ("mel", "bins", "f0_norm", "f0_raw"):
            np.save(os.path.join(shm, k + ".npy"), outs[k].cpu().numpy())
        jobs, cur, key = [], [], None
        for i in idx:
            m = mine[i]
            k2 = (m.spk_id, m.gender)
            if cur and (k2 != key or len(cur) >= per_task):
                jobs.append((shm, key[0], key[1], cur))
                cur = []
            key = k2
            cur.append((int(i), int(off[i]), int(off[i + 1]), int(fr[i]), int(fr[i + 1]), int(skips[i])))
        if cur:
            jobs.append((shm, key[0], key[1], cur))
        t1 = time.perf_counter()
        res = get_pool().map(_parity_job, jobs, chunksize=1)
        wall = time.perf_counter() - t1
    finally:
        shutil.rmtree(shm, ignore_errors=True)
    tot = dict(frames=0, same_bins=0, utts=0, utts_with_diff=0, voicing_mismatch=0, mel_over=0, cents_over=0,
               voiced_both=0, audio_s=0.0)
    hist = np.zeros(len(_HIST_EDGES) - 1, np.int64)
    mel_max = cents_max = f0n_max = 0.0
    worst, cells, band_max = [], [], np.zeros(80)
    for r in res:
        cells += r["cells"]
        band_max = np.maximum(band_max, r["band_max"])
        for k in tot:
            tot[k] += r[k]
        hist += r["hist"]
        mel_max, cents_max, f0n_max = max(mel_max, r["mel_max"]), max(cents_max, r["cents_max"]), max(f0n_max, r["f0_norm_max"])
        worst += r["worst"]
    cum = np.cumsum(hist) / max(1, hist.sum())

    def pct(q):                                    # upper edge of the histogram bin that holds quantile q
        return float(_HIST_EDGES[1:][int(np.searchsorted(cum, q))])

    frac = tot["same_bins"] / max(1, tot["frames"])
    worst.sort(key=lambda t: (-t[1], -t[2]))
    return {"utterances": int(tot["utts"]), "frames": int(tot["frames"]), "audio_s": tot["audio_s"],
            "scope": "every utterance of rank 0's shard" if which == "all" else "%d sampled utterances" % len(idx),
            "mel_max_abs": mel_max, "mel_p9999_abs_le": pct(0.9999), "mel_p50_abs_le": pct(0.5),
            "mel_values_over_1e-4": int(tot["mel_over"]), "mel_margin_to_gate": 1e-4 / max(mel_max, 1e-30),
            "identical_bins_frac": frac, "frames_with_different_bin": int(tot["frames"] - tot["same_bins"]),
            "utterances_with_any_different_bin": int(tot["utts_with_diff"]),
            "voicing_flag_mismatches": int(tot["voicing_mismatch"]),
            "f0_worst_cents_voiced_both": cents_max, "f0_frames_over_1_cent": int(tot["cents_over"]),
            "frames_voiced_in_both": int(tot["voiced_both"]), "f0_norm_max_abs": f0n_max,
            "mel_max_abs_per_band": [float("%.3g" % v) for v in band_max],
            "worst_mel_cells": [{"abs": a, "utt": b, "frame": c, "band": e, "ref": f, "got": g, "frames_in_utt": h}
                                for a, b, c, e, f, g, h in sorted(cells, reverse=True)[:12]],
            "worst_utterances": [{"index": a, "different_bins": b, "mel_max_abs": c} for a, b, c in worst[:5]],
            "oracle_wall_s": wall, "oracle_audio_s_per_s": tot["audio_s"] / max(wall, 1e-9), "setup_s": t1 - t0,
            "gates": "mel <= 1e-4 abs; bins + voicing identical on >= 99.9 % of frames; F0 <= 1 cent where voiced in both",
            "pass": bool(mel_max <= 1e-4 and frac >= 0.999)}


def cpu_baseline(mine, x, off, args):
    cores = pool_size()
    n_s = args.cpu_sample or min(len(mine), 200 * cores)
    pcm = {mine[i]: x[off[i]:off[i + 1]].cpu().numpy() for i in range(n_s)}
    jobs = make_cpu_jobs(mine[:n_s], pcm.__getitem__)
    secs, wall = cpu_time(jobs)
    return {"value": secs / wall, "unit": "audio-s/s", "cores": cores, "kind": "port",
            "sample": "first %d utterances of the corpus (%.0f audio-s), one speaker-atomic task per <=25 files, "
                      "multiprocessing.Pool(%d), BLAS threads 1; scipy filtfilt + numpy RandomState + pySTFT/mel "
                      "restatement + C RAPT restatement (pysptk/librosa absent)" % (n_s, secs, cores),
            "wall_s": wall}


# ---- reference arm ------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from speechsplit_b200.corpus import control_tracks, make_manifest, synth_batch
    os.environ.pop("WORLD_SIZE", None)          # rank 0 alone uses every host core
    cores = pool_size()
    metas = make_manifest(args.speakers, args.utts, seed=0)
    n_s = args.cpu_sample or min(len(metas), 100 * cores)
    sample = metas[:n_s]
    tracks = get_pool().map(control_tracks, sample, chunksize=16)
    pcm = []
    for s in range(0, n_s, 64):
        pcm += [p.numpy() for p in synth_batch(sample[s:s + 64], device="cpu", tracks=tracks[s:s + 64])]
    it = iter(pcm)
    jobs = make_cpu_jobs(sample, lambda m: next(it))
    for _ in range(min(args.warmup, 1)):
        cpu_time(jobs)
    secs_total, wall_total = 0.0, 0.0
    for _ in range(args.steps):
        secs, wall = cpu_time(jobs)
        secs_total += secs
        wall_total += wall
    value = secs_total / wall_total
    audio_s_total = float(sum(m.length for m in metas) / FS)
    desc = ("first %d utterances of the corpus (%.0f audio-s per step), speaker-atomic tasks, multiprocessing.Pool(%d), "
            "BLAS threads 1" % (n_s, secs_total / args.steps, cores))
    line = {"impl": "reference", "metric": "audio-sec/sec mel+F0", "value": value, "unit": "audio-s/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall_total / args.steps * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "VCTK-shaped synthetic corpus: %d speakers x %d utterances, %.0f audio-s, 16 kHz int16 PCM "
                                   "(BASELINE.json configs[1])" % (args.speakers, args.utts, audio_s_total),
                       "sample": desc},
            "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": cores, "kind": "port", "sample": desc,
                             "note": "the reference's own CPU path (scipy + numpy) with librosa.filters.mel and pysptk.rapt "
                                     "restated (neither is installed): oracle/"},
            "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def close_pool():
    global _POOL
    if _POOL is not None:
        _POOL.close()
        _POOL.join()
        _POOL = None


if __name__ == "__main__":
    a = parse()
    try:
        if a.impl == "reference":
            run_reference(a)
        else:
            run_ours(a)
    finally:
        close_pool()
