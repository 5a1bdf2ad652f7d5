"""GPU parity of the InterpLnr kernel through the C ABI: bit-exact against the
reference module's own outputs (tests/golden/interp_lnr.npz) and against the oracle restatement on
fresh seeded inputs; the module mirror consumes torch's generator like the reference."""
import os

import numpy as np
import pytest
import torch

from oracle.interp_lnr import interp_lnr as interp_ref
from speechsplit_b200.interp import InterpLnr

pytestmark = pytest.mark.gpu


def test_interp_lnr_golden(fe, golden_dir):
    z = np.load(os.path.join(golden_dir, "interp_lnr.npz"))
    mod = InterpLnr().cuda().train()
    for k in range(int(z["n"])):
        x = torch.from_numpy(z["x%d" % k]).cuda()
        y = mod.resample(x, torch.from_numpy(z["len_seq%d" % k]).cuda(), torch.from_numpy(z["scales%d" % k]).cuda(),
                         torch.from_numpy(z["len_seg%d" % k]).cuda())
        assert y.shape == z["y%d" % k].shape and y.dtype == torch.float32
        assert np.array_equal(y.cpu().numpy(), z["y%d" % k]), k


def test_interp_lnr_module_semantics(fe):
    """eval mode returns the input (model.py:382-383); training mode draws scales then segment lengths
    (the reference's order) and matches the oracle on those draws; no more than max_len_pad frames."""
    mod = InterpLnr().cuda()
    x = torch.rand((16, 192, 81), device="cuda")
    len_seq = torch.randint(64, 129, (16,), device="cuda")
    mod.eval()
    assert mod(x, len_seq) is x
    mod.train()
    torch.manual_seed(5)
    y = mod(x, len_seq)
    torch.manual_seed(5)
    scales, len_seg = mod.draw(16, x.device)
    ref = interp_ref(x.cpu().numpy(), len_seq.cpu().numpy(), scales.cpu().numpy(), len_seg.cpu().numpy().reshape(-1))
    assert np.array_equal(y.cpu().numpy(), ref)
    assert y.shape == (16, 192, 81)


def test_interp_lnr_other_geometry_and_errors(fe):
    mod = InterpLnr(max_len_seq=64, max_len_pad=96, min_len_seg=9, max_len_seg=16).cuda().train()
    x = torch.rand((5, 96, 3), device="cuda")
    len_seq = torch.tensor([96, 2, 64, 33, 10], device="cuda")
    scales, len_seg = mod.draw(5, x.device)
    y = mod.resample(x, len_seq, scales, len_seg)
    ref = interp_ref(x.cpu().numpy(), len_seq.cpu().numpy(), scales.cpu().numpy(), len_seg.cpu().numpy().reshape(-1),
                     max_len_seg=16, max_len_pad=96)
    assert np.array_equal(y.cpu().numpy(), ref)
    with pytest.raises(Exception):
        InterpLnr(max_len_pad=4096).cuda().train().resample(x, len_seq, scales, len_seg)
    with pytest.raises(RuntimeError):
        mod.resample(x.cpu(), len_seq, scales, len_seg)
