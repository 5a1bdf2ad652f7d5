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
    ap.add_argument("--parity-utts", type=int, default=12)
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


def get_pool():
    """Worker processes are forked once, BEFORE this process touches CUDA, and reused."""
    global _POOL
    if _POOL is None:
        world = int(os.environ.get("WORLD_SIZE", "1"))
        _POOL = mp.get_context("fork").Pool(max(1, (os.cpu_count() or 1) // max(1, world)))
    return _POOL


def pool_size():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    return max(1, (os.cpu_count() or 1) // max(1, world))


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
    from speechsplit_b200 import FrontEnd
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

    fe = FrontEnd(local)
    fix, fr = fe.plan(off)
    T = int(fr[-1])
    outs = dict(mel=torch.empty((T, 80), dtype=torch.float32, device=dev),
                f0_norm=torch.empty(T, dtype=torch.float32, device=dev),
                bins=torch.empty(T, dtype=torch.int64, device=dev),
                onehot=torch.empty((T, 257), dtype=torch.float32, device=dev))
    want = ("mel", "f0_norm", "bins", "onehot")

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
        del xh, ho

    collate = None
    if args.workload == "collate" and rank == 0:
        # configs[4]: batch 16 of 64..128-frame crops -> mel (16,192,80) clipped + one-hot (16,192,257)
        rng = np.random.default_rng(0)
        frs = np.diff(fr)
        cand = np.nonzero(frs > 130)[0]

        from speechsplit_b200.interp import InterpLnr
        interp = InterpLnr().to(dev).train()

        def crop_step():
            utt = rng.choice(cand, 16)
            ln = rng.integers(64, 129, 16)
            left = np.array([rng.integers(0, frs[u] - l) for u, l in zip(utt, ln)])
            melsp, pitch, onehot, bins = fe.collate(outs["mel"], outs["f0_norm"], fr, utt, left, ln, 192)
            # solver.py:160-161: the random-resampling augmentation of (mel, F0) - one kernel, no host sync
            x_intrp = interp(torch.cat((melsp, pitch), dim=-1), torch.from_numpy(ln).to(dev))
            return melsp, pitch, onehot, bins, x_intrp

        for _ in range(20):
            crop_step()
        torch.cuda.synchronize()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        c0.record()
        for _ in range(200):
            crop_step()
        c1.record()
        torch.cuda.synchronize()
        collate = {"steps_per_s_device": 200 / (c0.elapsed_time(c1) * 1e-3), "steps_per_s_wall": 200 / (time.perf_counter() - t0),
                   "batch": 16, "max_len_pad": 192,
                   "step": "crop + clip + pad + one-hot (data_loader.py:101-128, solver.py:162) + InterpLnr (model.py:380-436)"}

        # the same step through the reference-facing loader API (speechsplit_b200.data_loader.get_loader over
        # spmel / raptf0 NPY trees + train.pkl): one item per speaker = its first file, as data_loader.py:62-63
        import shutil
        from types import SimpleNamespace
        from speechsplit_b200.data_loader import get_loader, make_metadata
        tmp = tempfile.mkdtemp(prefix="ssfe_loader_")
        try:
            seen = set()
            for i, m in enumerate(mine):
                if m.spk in seen or frs[i] <= 130:
                    continue
                seen.add(m.spk)
                for sub, t in (("spmel", outs["mel"]), ("raptf0", outs["f0_norm"])):
                    os.makedirs(os.path.join(tmp, sub, m.spk), exist_ok=True)
                    np.save(os.path.join(tmp, sub, m.spk, "%s_001.npy" % m.spk), t[fr[i]:fr[i + 1]].cpu().numpy(),
                            allow_pickle=False)
            make_metadata(os.path.join(tmp, "spmel"), verbose=False)
            n_batches = 220
            hp = SimpleNamespace(root_dir=os.path.join(tmp, "spmel"), feat_dir=os.path.join(tmp, "raptf0"), mode="train",
                                 batch_size=16, shuffle=True, num_workers=0, samplier=-(-16 * n_batches // len(seen)),
                                 min_len_seq=64, max_len_seq=128, max_len_pad=192)
            for mode in ("reference", "batched"):
                loader = get_loader(hp, frontend=fe, want_onehot=True, draws=mode)
                it = iter(loader)

                def loader_step():
                    melsp, emb, pitch, len_org = next(it)                       # solver.py:142
                    return interp(torch.cat((melsp, pitch), dim=-1), len_org)   # solver.py:160-161

                for _ in range(20):
                    loader_step()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                c0.record()
                for _ in range(n_batches - 20):
                    loader_step()
                c1.record()
                torch.cuda.synchronize()
                key = "loader" if mode == "reference" else "loader_batched_draws"
                collate[key + "_steps_per_s_device"] = (n_batches - 20) / (c0.elapsed_time(c1) * 1e-3)
                collate[key + "_steps_per_s_wall"] = (n_batches - 20) / (time.perf_counter() - t0)
            collate["loader"] = ("speechsplit_b200.data_loader.get_loader, %d speakers, features resident in HBM; "
                                 "draws in the reference's order (two np.random.randint calls per item) / batched "
                                 "(two per batch)" % len(seen))
            if not args.no_cpu_baseline:
                # the reference's collator loop + quantisation on one host core (its loader default is num_workers=0)
                from oracle import collate_ref, ref_pipeline
                items = [tuple(loader.dataset[i]) for i in range(len(loader.dataset))]
                t0, nb = time.perf_counter(), 0
                while time.perf_counter() - t0 < 3.0:
                    bt = [items[j] for j in rng.integers(0, len(items), 16)]
                    _, _, pitch_ref, _ = collate_ref.collate(bt, 64, 128, 192)
                    ref_pipeline.quantize_f0_numpy(pitch_ref.reshape(-1))
                    nb += 1
                collate["cpu_steps_per_s"] = nb / (time.perf_counter() - t0)
                collate["cpu_step"] = "oracle.collate_ref (data_loader.py:101-128) + quantize_f0_numpy, 1 core, no InterpLnr"
        finally:
            shutil.rmtree(tmp, ignore_errors=True)

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
    roofline = {"kernel": "stft_mel_kernel<0> (fused STFT->mel->dB->normalise)", "bound": "hbm",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s",
                "traffic": None, "frames_per_launch": T, "bytes_per_frame": BYTES_PER_FRAME,
                "launch_ms": stft_ms, "fft_tflops_nominal": T * FLOP_PER_FRAME_FFT / (stft_ms * 1e-3) / 1e12,
                "note": "issue / shared-memory bound, not HBM bound (DESIGN.md 5): 1344 B/frame vs ~790 warp instructions and 322 shared-memory wavefronts"}
    # the binding roof of this kernel is FP32 issue, so that fraction is reported beside the HBM one
    # (70.6 TFLOP/s = the FMA rate measured with profiles/microbench/fp64_bench.cu on this pool's B200s)
    roofline["fp32_peak_tflops"] = 70.6
    roofline["fft_frac_of_fp32_peak"] = roofline["fft_tflops_nominal"] / 70.6
    tr = os.path.join(ROOT, "profiles", "stft_traffic.json")
    if os.path.exists(tr):
        try:
            t = json.load(open(tr))
            roofline["traffic"] = t["dram_bytes_per_frame"] * T
            roofline["traffic_source"] = t.get("source")
        except Exception:
            pass
    stages = {k: v / args.steps for k, v in stage_acc.items()}

    # ---- parity gate on a sample of this very corpus -------------------------------------------------
    parity = parity_check(mine, skips, x, off, fr, outs, args.parity_utts)

    # ---- CPU baseline on a bounded sample ---------------------------------------------------------------
    cpu = None
    if not args.no_cpu_baseline:
        cpu = cpu_baseline(mine, x, off, args)

    line = {"metric": "audio-sec/sec mel+F0", "value": value, "unit": "audio-s/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32 (STFT/mel), f64 (filtfilt), f32+f64 (RAPT)",
            "data": "synthetic",
            "config": {"workload": "VCTK-shaped synthetic corpus: %d speakers x %d utterances, %.0f audio-s, 16 kHz int16 PCM "
                                   "(BASELINE.json configs[1]); at N>1 the same corpus cut into N consecutive runs of equal sample count (configs[2])"
                                   % (args.speakers, args.utts, audio_s_total),
                       "utterances": len(metas), "frames": int(T) if world == 1 else None,
                       "outputs": "mel f32 [T,80], f0_norm f32 [T], bins i64 [T], one-hot f32 [T,257]",
                       "l2": "inputs per step (%.1f GB PCM on rank 0) exceed the 126 MB L2; no flush needed" % (x.numel() * 2 / 1e9),
                       "collective": "none on the data path"},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
            "cpu_baseline": cpu, "stage_ms": stages, "parity": parity}
    if args.workload != "vctk":
        line["config"]["workload"] = {"longform": "BASELINE configs[3]: 256 x 60.000 s utterances (4 speakers x 64)",
                                      "single": "BASELINE configs[0]: one 3 s male utterance (p226)",
                                      "collate": line["config"]["workload"] + " + configs[4] collator"}[args.workload]
    if collate:
        line["collate"] = collate
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def parity_check(mine, skips, x, off, fr, outs, n_check):
    """Oracle vs GPU on a few utterances of the benchmarked corpus (the gates of SURVEY.md 8(d))."""
    from numpy.random import RandomState
    from oracle import ref_pipeline as rp
    if n_check <= 0:
        return None
    idx = np.linspace(0, len(mine) - 1, n_check).astype(int)
    worst_mel, same, tot, cents = 0.0, 0, 0, 0.0
    for i in idx:
        m = mine[i]
        p = x[off[i]:off[i + 1]].cpu().numpy()
        prng = RandomState(m.spk_id)
        if int(skips[i]):
            prng.rand(int(skips[i]))             # advance the speaker's stream to this file
        S, f0n = rp.extract_utterance(p.astype(np.float64) / 32768.0, m.gender, prng)
        gm = outs["mel"][fr[i]:fr[i + 1]].cpu().numpy()
        gb = outs["bins"][fr[i]:fr[i + 1]].cpu().numpy()
        gf = outs["f0_norm"][fr[i]:fr[i + 1]].cpu().numpy()
        assert gm.shape == S.shape
        worst_mel = max(worst_mel, float(np.abs(gm - S).max()))
        rb = rp.quantize_f0_numpy(f0n)[1]
        same += int((gb == rb).sum())
        tot += rb.size
        both = (gf > 0) & (f0n > 0)
        if both.any():
            cents = max(cents, float(np.abs(gf[both] - f0n[both]).max()))
    return {"utterances": int(len(idx)), "mel_max_abs": worst_mel, "identical_bins_frac": same / max(tot, 1),
            "f0_norm_max_abs": cents, "pass": bool(worst_mel <= 1e-4 and same >= 0.999 * tot)}


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
