"""GPU diagnostic: dump PCM + GPU results of the utterances with the worst mel deviations (bench corpus)."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "."))
import bench
sys.argv = ["bench.py"]
args = bench.parse()
metas, mine, skips, tracks = bench.build_shard(args, 0, 1)
import torch
from speechsplit_b200 import FrontEnd, FrontEndConfig
dev = torch.device("cuda", 0)
x, off = bench.synth_on_gpu(mine, tracks, dev)
bench.close_pool()
par = json.loads(open(os.path.join(os.environ.get("GRAFT_REPO_ROOT", "."), "scratch", "r2_par_scan.json")).read().strip().splitlines()[-1])["parity"]
utts = sorted({c["utt"] for c in par["worst_mel_cells"]})
fe, fes = FrontEnd(0), FrontEnd(0, FrontEndConfig(filtfilt_mode=1))
out = {}
for i in utts:
    m = mine[i]
    p = x[off[i]:off[i + 1]].clone()
    lo, hi = ([50.0], [250.0]) if m.gender == "M" else ([100.0], [600.0])
    a = fe.extract(p, [0, len(p)], lo, hi, [m.spk_id], [int(skips[i])], want=("mel", "f0_norm", "wav"))      # raw-dither path
    b = fe.extract(p, [0, len(p)], lo, hi, [m.spk_id], [int(skips[i])], want=("mel", "f0_norm"))             # production path
    c = fes.extract(p, [0, len(p)], lo, hi, [m.spk_id], [int(skips[i])], want=("mel", "f0_norm", "wav64"))
    y, _ = fe.filtfilt(p, [0, len(p)])
    ys, _ = fes.filtfilt(p, [0, len(p)])
    out["pcm%d" % i] = p.cpu().numpy()
    out["meta%d" % i] = np.array([m.spk_id, int(skips[i]), 1 if m.gender == "M" else 0], np.int64)
    out["mel_raw%d" % i] = a["mel"].cpu().numpy()
    out["wav_raw%d" % i] = a["wav"].cpu().numpy()
    out["mel_prod%d" % i] = b["mel"].cpu().numpy()
    out["mel_seq%d" % i] = c["mel"].cpu().numpy()
    out["wav64_seq%d" % i] = c["wav64"].cpu().numpy()
    out["y_scan%d" % i] = y.cpu().numpy()
    out["y_seq%d" % i] = ys.cpu().numpy()
out["utts"] = np.array(utts)
np.savez_compressed("gpurun_out/r2_diag_mel.npz", **out)
print("dumped", utts)
