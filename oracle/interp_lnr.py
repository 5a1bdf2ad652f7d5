"""oracle/interp_lnr.py - TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

CPU restatement of ``InterpLnr.forward`` in training mode (reference model.py:380-436), with the
random draws (``scales`` :392-393, ``len_seg`` :401-404) passed in.  Arithmetic is float32 in torch's
evaluation order, so the result is bit-identical to the reference's for the same draws; pinned by
tests/golden/interp_lnr.npz, which the reference module itself produced (tests/golden/make_golden.py).
"""
import numpy as np


def interp_lnr(x, len_seq, scales, len_seg, max_len_seg=32, max_len_pad=192):
    """x (B, T, C) float32, len_seq (B,) int, scales (B * S,) float32, len_seg (B * S,) int ->
    (B, max_len_pad, C) float32."""
    x = np.asarray(x, np.float32)
    B, T, C = x.shape
    scales = np.asarray(scales, np.float32).reshape(B, -1)
    len_seg = np.asarray(len_seg, np.int64).reshape(B, -1)
    S = scales.shape[1]
    out = np.zeros((B, max_len_pad, C), np.float32)
    idx = np.arange(2 * max_len_seg, dtype=np.float32)                       # :388-389
    for b in range(B):
        rows = []
        offset = 0
        for s in range(S):
            idx_scaled = (idx / scales[b, s]).astype(np.float32)              # :397
            fl = np.floor(idx_scaled)                                         # :398
            lam = (idx_scaled - fl).astype(np.float32)                        # :399
            org = (fl + np.float32(offset)).astype(np.float32)                # :412
            mask = (fl < np.float32(len_seg[b, s] - 1)) & (org < np.float32(len_seq[b] - 1))   # :405, :414-417
            i0 = org[mask].astype(np.int64)                                   # :424
            lm = lam[mask][:, None]
            y = ((np.float32(1) - lm) * x[b, i0, :]).astype(np.float32) + (lm * x[b, i0 + 1, :]).astype(np.float32)   # :427
            rows.append(y.astype(np.float32))
            offset += int(len_seg[b, s])                                      # :408-410
        seq = np.concatenate(rows, axis=0)[:max_len_pad]                      # :366-377
        out[b, :seq.shape[0]] = seq
    return out
