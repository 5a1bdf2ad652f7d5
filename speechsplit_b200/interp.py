"""InterpLnr (reference model.py:355-436) with the training-mode forward as one CUDA kernel.

Same constructor and call as the reference module (``InterpLnr(hparams)``; ``forward(x, len_seq)``
returns ``x`` unchanged in eval mode).  The random draws are made with the reference's torch calls in
the reference's order, so a seeded run consumes the generator exactly as the reference does; the
resampling itself - masks, gather, linear interpolation, per-item concatenation, zero padding - is
``ssfe_interp_lnr`` (csrc/interp.cu), without the reference's host sync (``counts.tolist()``, :432).
"""
import torch

from . import _lib as L
from .frontend import default_frontend


class InterpLnr(torch.nn.Module):
    def __init__(self, hparams=None, *, max_len_seq=128, max_len_pad=192, min_len_seg=19, max_len_seg=32):
        super().__init__()
        g = (lambda k, d: getattr(hparams, k, d)) if hparams is not None else (lambda k, d: d)
        self.max_len_seq = g("max_len_seq", max_len_seq)
        self.max_len_pad = g("max_len_pad", max_len_pad)
        self.min_len_seg = g("min_len_seg", min_len_seg)
        self.max_len_seg = g("max_len_seg", max_len_seg)
        self.max_num_seg = self.max_len_seq // self.min_len_seg + 1          # model.py:364

    def draw(self, batch_size, device):
        """The two random tensors of model.py:392-393 and :401-404, in that order."""
        scales = torch.rand(batch_size * self.max_num_seg, device=device) + 0.5
        len_seg = torch.randint(low=self.min_len_seg, high=self.max_len_seg,
                                size=(batch_size * self.max_num_seg, 1), device=device)
        return scales, len_seg

    def resample(self, x, len_seq, scales, len_seg):
        if not x.is_cuda:
            raise RuntimeError("InterpLnr: the training-mode path is a CUDA kernel; there is no CPU fallback")
        fe = default_frontend(x.device.index)
        x = x.contiguous().float()
        B, T, C = x.shape
        len_seq = len_seq.to(device=x.device, dtype=torch.int64).contiguous()
        scales = scales.to(device=x.device, dtype=torch.float32).contiguous().view(-1)
        len_seg = len_seg.to(device=x.device, dtype=torch.int64).contiguous().view(-1)
        out = torch.empty((B, self.max_len_pad, C), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            fe._bind_stream()
            fe._check(fe.lib.ssfe_interp_lnr(fe._h, L.vp(x.data_ptr()), B, T, C, L.vp(len_seq.data_ptr()),
                                             L.vp(scales.data_ptr()), L.vp(len_seg.data_ptr()), self.max_num_seg,
                                             self.max_len_seg, self.max_len_pad, L.vp(out.data_ptr())))
        return out

    def forward(self, x, len_seq):
        if not self.training:                                                # model.py:382-383
            return x
        scales, len_seg = self.draw(x.size(0), x.device)
        return self.resample(x, len_seq, scales, len_seg)
