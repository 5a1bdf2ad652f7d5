"""oracle/torch_eager_ref.py - TEST / BENCH INFRASTRUCTURE ONLY (see oracle/__init__.py).

What the reference itself runs ON THE GPU, in torch eager, for the two steps between its collator and
its model (solver.py:160-163): ``InterpLnr.forward`` in training mode (model.py:380-436) and
``utils.quantize_f0_torch`` (utils.py:62-74).  Restated op for op - same tensor ops in the same order,
including the boolean-index gathers, the ``repeat_interleave`` calls and the ``counts.tolist()`` host
sync of model.py:432 - so that ``bench.py`` can time the reference's own launch pattern beside the
one-kernel versions (``ssfe_interp_lnr``, ``ssfe_collate``) on the same device, and the tests can compare
results.  PINNED: tests/test_oracle.py checks it against tests/golden/interp_lnr.npz (produced by the
reference's own module) and against utils_kat.npz for the quantiser.  Never imported by the product.
"""
import torch
import torch.nn.functional as F


def quantize_f0_eager(x, num_bins=256):
    """utils.py:62-74: (B, T) float -> (one-hot (B, T, num_bins + 1) f32, bins (B, T) i64)."""
    n_batch = x.size(0)                                   # :64
    flat = x.view(-1).clone()                             # :65
    unvoiced = flat <= 0                                  # :66
    flat[unvoiced] = 0                                    # :67
    assert (flat >= 0).all() and (flat <= 1).all()        # :68
    flat = torch.round(flat * (num_bins - 1))             # :69
    flat = flat + 1                                       # :70
    flat[unvoiced] = 0                                    # :71
    enc = torch.zeros((flat.size(0), num_bins + 1), device=flat.device)     # :72
    enc[torch.arange(flat.size(0)), flat.long()] = 1                        # :73
    return enc.view(n_batch, -1, num_bins + 1), flat.view(n_batch, -1).long()   # :74


def interp_lnr_eager(x, len_seq, max_len_seq=128, max_len_pad=192, min_len_seg=19, max_len_seg=32, draws=None):
    """model.py:380-436 (training mode).  ``draws`` = (scales, len_seg) replaces the two random calls
    (:392-393, :401-404) so results can be compared; None draws them as the reference does."""
    dev = x.device
    n_batch = x.size(0)
    n_seg = max_len_seq // min_len_seg + 1                                   # :365
    idx = torch.arange(max_len_seg * 2, device=dev).unsqueeze(0).expand(n_batch * n_seg, -1)   # :388-389
    if draws is None:
        scales = torch.rand(n_batch * n_seg, device=dev) + 0.5               # :392-393
        len_seg = torch.randint(low=min_len_seg, high=max_len_seg, size=(n_batch * n_seg, 1), device=dev)   # :401-404
    else:
        scales, len_seg = draws
        scales = scales.to(dev).view(-1)
        len_seg = len_seg.to(dev).view(-1, 1)
    scaled = idx / scales.unsqueeze(-1)                                      # :395
    scaled_fl = torch.floor(scaled)                                          # :396
    lam = scaled - scaled_fl                                                 # :397
    in_seg = scaled_fl < (len_seg - 1)                                       # :407
    offset = len_seg.view(n_batch, -1).cumsum(dim=-1)                        # :409
    offset = F.pad(offset[:, :-1], (1, 0), value=0).view(-1, 1)              # :411
    org = scaled_fl + offset                                                 # :413
    len_rp = torch.repeat_interleave(len_seq, n_seg)                         # :415
    in_seq = org < (len_rp - 1).unsqueeze(-1)                                # :416
    keep = in_seg & in_seq                                                   # :418
    counts = keep.sum(dim=-1).view(n_batch, -1).sum(dim=-1)                  # :420
    row = torch.repeat_interleave(torch.arange(n_batch, device=dev), counts)  # :422-423
    i_fl = org[keep].long()                                                  # :425
    i_cl = i_fl + 1                                                          # :426
    y_fl = x[row, i_fl, :]                                                   # :428
    y_cl = x[row, i_cl, :]                                                   # :429
    lam_k = lam[keep].unsqueeze(-1)                                          # :430
    y = (1 - lam_k) * y_fl + lam_k * y_cl                                    # :432
    pieces = torch.split(y, counts.tolist(), dim=0)                          # :434 (host sync)
    out = x.new_zeros((len(pieces), max_len_pad, x.size(-1)))                # :368-370
    for i, t in enumerate(pieces):                                           # :372-374
        n = t.size(0)
        out[i, :n, :] = t[:max_len_pad]
    return out
