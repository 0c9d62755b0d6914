"""Kernel-level parity of the bandwidth kernels (through the C ABI) vs plain PyTorch fp32 on the same inputs."""
import math
import os

import pytest
import torch
import torch.nn.functional as F

from oracle import functional as OF

pytestmark = pytest.mark.gpu


def _bf(t):
    return t.to(torch.bfloat16).float()


@pytest.mark.parametrize("variant", ["tiled", "direct"])
@pytest.mark.parametrize("C,H,W,K,S", [(64, 21, 45, 3, 1), (24 * 8, 20, 33, 3, 2), (40, 17, 29, 5, 1), (288, 9, 14, 5, 2),
                                        (480, 12, 43, 5, 1), (24, 19, 30, 3, 2), (16, 33, 37, 5, 2),
                                        (384, 70, 150, 5, 1)])
def test_dwconv_silu_and_squeeze(C, H, W, K, S, variant):
    """depthwise conv + bias + SiLU (TF-SAME padding) and the squeeze partial sums"""
    from occdepth_b200 import _lib
    L = _lib.lib()
    g = torch.Generator().manual_seed(C + K)
    B = 2
    x = _bf(torch.randn(B, C, H, W, generator=g))
    w = torch.randn(C, 1, K, K, generator=g) / K
    b = torch.randn(C, generator=g)
    OH, OW = math.ceil(H / S), math.ceil(W / S)
    ph = max((OH - 1) * S + K - H, 0)
    pw = max((OW - 1) * S + K - W, 0)
    ref = F.silu(F.conv2d(F.pad(x, [pw // 2, pw - pw // 2, ph // 2, ph - ph // 2]), w, b, S, 0, 1, C))
    xc = x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).cuda()
    y = torch.empty(B, OH, OW, C, dtype=torch.bfloat16, device="cuda")
    wk = w.reshape(C, K * K).t().contiguous().cuda()
    pool = torch.zeros(B, C, dtype=torch.int64, device="cuda")
    fn = L.occd_dwconv2d_tiled_fwd if variant == "tiled" else L.occd_dwconv2d_fwd
    rc = fn(xc.data_ptr(), wk.data_ptr(), b.cuda().data_ptr(), y.data_ptr(), pool.data_ptr(), B, H, W,
            OH, OW, C, C, C, K, S, ph // 2, pw // 2, _lib.ACT_SILU, _lib.stream_ptr())
    assert rc == 0
    got = y.float().cpu().permute(0, 3, 1, 2)
    assert float((got - ref).abs().max()) <= 2 ** -7 * float(ref.abs().max())
    sums = pool.cpu().double().div(2 ** 24).float()
    assert float((sums - got.sum((2, 3))).abs().max()) <= 1e-3 * float(got.sum((2, 3)).abs().max())


@pytest.mark.parametrize("variant", ["wide", "strip"])
def test_se_gate_fold(variant):
    from occdepth_b200 import _lib
    L = _lib.lib()
    if variant == "strip" and os.environ.get("OCCD_EXPERIMENTAL") != "1":
        pytest.skip("strip fold: CPU-emulation tested (tests/test_se_fold_host.py); first GPU run is opt-in")
    g = torch.Generator().manual_seed(0)
    C, R, rows = 192, 12, 48
    Kp = 192
    pool = (torch.randn(1, C, generator=g) * 50 * 2 ** 24).round().to(torch.int64)
    w1, b1 = torch.randn(R, C, generator=g) / C ** 0.5, torch.randn(R, generator=g)
    w2, b2 = torch.randn(C, R, generator=g) / R ** 0.5, torch.randn(C, generator=g)
    master = torch.randn(rows, Kp, generator=g)
    hw = 77.0
    gate = torch.sigmoid(F.linear(F.silu(F.linear(pool.double().div(2 ** 24).float() / hw, w1, b1)), w2, b2))
    want = (master * gate).to(torch.bfloat16).float()
    out = torch.zeros(rows, Kp, dtype=torch.bfloat16, device="cuda")
    hid = torch.zeros(R, device="cuda")
    d = lambda t: t.contiguous().cuda()
    bufs = [d(pool), d(w1), d(b1), d(w2.t()), d(b2), d(master)]
    fn = L.occd_se_gate_fold_strip_fwd if variant == "strip" else L.occd_se_gate_fold_fwd
    rc = fn(bufs[0].data_ptr(), 1.0 / hw, bufs[1].data_ptr(), bufs[2].data_ptr(), bufs[3].data_ptr(),
            bufs[4].data_ptr(), hid.data_ptr(), bufs[5].data_ptr(), out.data_ptr(), 1, C, R, rows, Kp,
            _lib.stream_ptr())
    assert rc == 0
    assert float((out.float().cpu() - want).abs().max()) <= 2 ** -7 * float(want.abs().max())


@pytest.mark.parametrize("variant", ["flat", "rows"])
def test_bilinear_align_corners(variant):
    from occdepth_b200 import _lib
    L = _lib.lib()
    if variant == "rows" and os.environ.get("OCCD_EXPERIMENTAL") != "1":
        pytest.skip("rows resize: CPU-emulation tested (tests/test_upsample_host.py); first GPU run is opt-in")
    fn = L.occd_upsample_bilinear_rows if variant == "rows" else L.occd_upsample_bilinear_ac
    g = torch.Generator().manual_seed(0)
    B, C, h, w, OH, OW = 2, 24, 7, 9, 12, 22
    x = _bf(torch.randn(B, C, h, w, generator=g))
    ref = F.interpolate(x, size=(OH, OW), mode="bilinear", align_corners=True)
    xc = x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).cuda()
    y = torch.zeros(B, OH, OW, C, dtype=torch.bfloat16, device="cuda")
    assert fn(xc.data_ptr(), y.data_ptr(), B, h, w, OH, OW, C, C, 0, C, 0, _lib.stream_ptr()) == 0
    got = y.float().cpu().permute(0, 3, 1, 2)
    assert float((got - ref).abs().max()) <= 2 ** -7 * float(ref.abs().max())


def test_virtual_view_kernel():
    from occdepth_b200 import _lib
    L = _lib.lib()
    g = torch.Generator().manual_seed(0)
    B, C, h, w, H, W = 1, 16, 8, 16, 32, 64
    x = _bf(torch.randn(B, C, h, w, generator=g))
    depth = torch.rand(1, 1, H, W, generator=g) * 7 + 0.5
    depth[0, 0, 5, 9] = 0.0
    bf, s = 6.0, 4
    ref = OF.virtual_view(x, depth, s, bf)
    xc = x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).cuda()
    y = torch.zeros(B, h, w, C, dtype=torch.bfloat16, device="cuda")
    dd = depth[0, 0].contiguous().cuda()
    assert L.occd_virtual_view_fwd(xc.data_ptr(), y.data_ptr(), dd.data_ptr(), B, h, w, C, C, C, H, W, bf / s,
                                   _lib.stream_ptr()) == 0
    got = y.float().cpu().permute(0, 3, 1, 2)
    assert float((got - ref).abs().max()) <= 2 ** -6 * float(ref.abs().max())


def test_softmax_planar_and_fc():
    from occdepth_b200 import _lib
    L = _lib.lib()
    g = torch.Generator().manual_seed(0)
    x = torch.randn(3, 37, 50, generator=g) * 3
    out = torch.empty(3, 37, 50, device="cuda")
    assert L.occd_softmax_planar(x.cuda().data_ptr(), out.data_ptr(), 3, 37, 50, _lib.stream_ptr()) == 0
    assert float((out.cpu() - x.softmax(1)).abs().max()) <= 1e-6
    inp, w, b = torch.randn(4, 33, generator=g), torch.randn(20, 33, generator=g), torch.randn(20, generator=g)
    o = torch.empty(4, 20, device="cuda")
    assert L.occd_fc_fwd(inp.cuda().data_ptr(), w.cuda().data_ptr(), b.cuda().data_ptr(), o.data_ptr(), 4, 33, 20,
                         _lib.ACT_RELU, _lib.stream_ptr()) == 0
    assert float((o.cpu() - F.relu(F.linear(inp, w, b))).abs().max()) <= 1e-5

