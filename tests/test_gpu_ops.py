"""Kernel-level parity of the bandwidth kernels (through the C ABI) vs plain PyTorch fp32 on the same inputs."""
import math

import pytest
import torch
import torch.nn.functional as F

from oracle import functional as OF

pytestmark = pytest.mark.gpu


def tf32r(t):
    return ((t.contiguous().view(torch.int32) + 0x1000) & -0x2000).view(torch.float32)


# (torch dtype, C-ABI dtype code, operand rounding, one-store tolerance relative to max |ref| = 2x the rounding step)
MODES = {"tf32": (torch.float32, 0, tf32r, 2.0 ** -10), "bf16": (torch.bfloat16, 1, lambda t: t.to(torch.bfloat16).float(),
                                                               2.0 ** -8)}
PREC = pytest.mark.parametrize("precision", sorted(MODES))


@PREC
@pytest.mark.parametrize("variant", ["tiled", "direct"])
@pytest.mark.parametrize("C,H,W,K,S", [(64, 21, 45, 3, 1), (24 * 8, 20, 33, 3, 2), (40, 17, 29, 5, 1), (288, 9, 14, 5, 2),
                                        (480, 12, 43, 5, 1), (24, 19, 30, 3, 2), (16, 33, 37, 5, 2),
                                        (384, 70, 150, 5, 1)])
def test_dwconv_silu_and_squeeze(C, H, W, K, S, variant, precision):
    """depthwise conv + bias + SiLU (TF-SAME padding) and the squeeze partial sums"""
    from occdepth_b200 import _lib
    L = _lib.lib()
    tdt, code, rnd, tol = MODES[precision]
    g = torch.Generator().manual_seed(C + K)
    B = 2
    x = rnd(torch.randn(B, C, H, W, generator=g))
    w = torch.randn(C, 1, K, K, generator=g) / K
    b = torch.randn(C, generator=g)
    OH, OW = math.ceil(H / S), math.ceil(W / S)
    ph = max((OH - 1) * S + K - H, 0)
    pw = max((OW - 1) * S + K - W, 0)
    ref = F.silu(F.conv2d(F.pad(x, [pw // 2, pw - pw // 2, ph // 2, ph - ph // 2]), w, b, S, 0, 1, C))
    xc = x.permute(0, 2, 3, 1).contiguous().to(tdt).cuda()
    y = torch.empty(B, OH, OW, C, dtype=tdt, device="cuda")
    wk = w.reshape(C, K * K).t().contiguous().cuda()
    pool = torch.zeros(B, C, dtype=torch.int64, device="cuda")
    fn = L.occd_dwconv2d_tiled_fwd if variant == "tiled" else L.occd_dwconv2d_fwd
    rc = fn(xc.data_ptr(), wk.data_ptr(), b.cuda().data_ptr(), y.data_ptr(), pool.data_ptr(), code, B, H, W,
            OH, OW, C, C, C, K, S, ph // 2, pw // 2, _lib.ACT_SILU, _lib.stream_ptr())
    assert rc == 0
    got = y.float().cpu().permute(0, 3, 1, 2)
    assert float((got - ref).abs().max()) <= tol * float(ref.abs().max())
    if precision == "tf32":
        assert torch.equal(got, tf32r(got))          # every stored activation is a TF32 value
    sums = pool.cpu().double().div(2 ** 24).float()
    assert float((sums - got.sum((2, 3))).abs().max()) <= 1e-3 * float(got.sum((2, 3)).abs().max())


@PREC
def test_se_gate_fold(precision):
    from occdepth_b200 import _lib
    L = _lib.lib()
    tdt, code, rnd, tol = MODES[precision]
    g = torch.Generator().manual_seed(0)
    C, R, rows = 192, 12, 48
    Kp = 192
    pool = (torch.randn(1, C, generator=g) * 50 * 2 ** 24).round().to(torch.int64)
    w1, b1 = torch.randn(R, C, generator=g) / C ** 0.5, torch.randn(R, generator=g)
    w2, b2 = torch.randn(C, R, generator=g) / R ** 0.5, torch.randn(C, generator=g)
    master = torch.randn(rows, Kp, generator=g)
    hw = 77.0
    gate = torch.sigmoid(F.linear(F.silu(F.linear(pool.double().div(2 ** 24).float() / hw, w1, b1)), w2, b2))
    want = rnd(master * gate)
    out = torch.zeros(rows, Kp, dtype=tdt, device="cuda")
    hid = torch.zeros(R, device="cuda")
    d = lambda t: t.contiguous().cuda()
    bufs = [d(pool), d(w1), d(b1), d(w2.t()), d(b2), d(master)]
    rc = L.occd_se_gate_fold_fwd(bufs[0].data_ptr(), 1.0 / hw, bufs[1].data_ptr(), bufs[2].data_ptr(),
                                 bufs[3].data_ptr(), bufs[4].data_ptr(), hid.data_ptr(), bufs[5].data_ptr(),
                                 out.data_ptr(), code, 1, C, R, rows, Kp, _lib.stream_ptr())
    assert rc == 0
    assert float((out.float().cpu() - want).abs().max()) <= tol * float(want.abs().max())
    assert int(bufs[0].abs().max()) == 0            # the squeeze sums are cleared for the next forward


@PREC
def test_bilinear_align_corners(precision):
    from occdepth_b200 import _lib
    L = _lib.lib()
    tdt, code, rnd, tol = MODES[precision]
    g = torch.Generator().manual_seed(0)
    B, C, h, w, OH, OW = 2, 24, 7, 9, 12, 22
    x = rnd(torch.randn(B, C, h, w, generator=g))
    ref = F.interpolate(x, size=(OH, OW), mode="bilinear", align_corners=True)
    xc = x.permute(0, 2, 3, 1).contiguous().to(tdt).cuda()
    y = torch.zeros(B, OH, OW, C, dtype=tdt, device="cuda")
    assert L.occd_upsample_bilinear_ac(xc.data_ptr(), y.data_ptr(), code, B, h, w, OH, OW, C, C, 0, C, 0,
                                       _lib.stream_ptr()) == 0
    got = y.float().cpu().permute(0, 3, 1, 2)
    assert float((got - ref).abs().max()) <= tol * float(ref.abs().max())


@PREC
def test_virtual_view_kernel(precision):
    from occdepth_b200 import _lib
    L = _lib.lib()
    tdt, code, rnd, tol = MODES[precision]
    g = torch.Generator().manual_seed(0)
    B, C, h, w, H, W = 1, 16, 8, 16, 32, 64
    x = rnd(torch.randn(B, C, h, w, generator=g))
    depth = torch.rand(1, 1, H, W, generator=g) * 7 + 0.5
    depth[0, 0, 5, 9] = 0.0
    bf, s = 6.0, 4
    ref = OF.virtual_view(x, depth, s, bf)
    xc = x.permute(0, 2, 3, 1).contiguous().to(tdt).cuda()
    y = torch.zeros(B, h, w, C, dtype=tdt, device="cuda")
    dd = depth[0, 0].contiguous().cuda()
    assert L.occd_virtual_view_fwd(xc.data_ptr(), y.data_ptr(), dd.data_ptr(), code, B, h, w, C, C, C, H, W, bf / s,
                                   _lib.stream_ptr()) == 0
    got = y.float().cpu().permute(0, 3, 1, 2)
    # the sampling coordinate itself is computed in fp32 on both sides but with different operation order: a
    # coordinate differing in the last bit moves a bilinear weight by ~1e-6 of |x|
    assert float((got - ref).abs().max()) <= (tol + 1e-4) * float(ref.abs().max())


@PREC
def test_small_channels_last_helpers(precision):
    """softmax -> channel window, channel-window copy, channels-last -> K-major weight transpose, channel scale"""
    from occdepth_b200 import _lib
    L = _lib.lib()
    tdt, code, rnd, tol = MODES[precision]
    g = torch.Generator().manual_seed(3)
    st = _lib.stream_ptr()
    B, S, cs = 2, 77, 16
    x = torch.randn(B, 2, S, generator=g) * 2
    out = torch.zeros(B, S, cs, dtype=tdt, device="cuda")
    assert L.occd_softmax_planar_to_cl(x.cuda().data_ptr(), out.data_ptr(), code, B, 2, S, cs, 8, st) == 0
    got = out.float().cpu()
    assert float((got[..., 8:10] - x.softmax(1).permute(0, 2, 1)).abs().max()) <= tol
    assert float(got[..., :8].abs().max()) == 0 and float(got[..., 10:].abs().max()) == 0
    src = rnd(torch.randn(B * S, 24, generator=g)).to(tdt).cuda()
    dst = torch.zeros(B * S, 40, dtype=tdt, device="cuda")
    assert L.occd_copy_channels(src.data_ptr(), dst.data_ptr(), code, B * S, 16, 24, 8, 40, 16, st) == 0
    assert torch.equal(dst[:, 16:32], src[:, 8:24]) and float(dst[:, :16].float().abs().max()) == 0
    P, C, ldo = 45, 20, 48
    a = rnd(torch.randn(B, P, 24, generator=g)).to(tdt).cuda()
    wt = torch.zeros(B, 32, ldo, dtype=tdt, device="cuda")
    assert L.occd_cl_transpose(a.data_ptr(), wt.data_ptr(), code, B, P, C, 24, 0, ldo, 32 * ldo, st) == 0
    assert torch.equal(wt[:, :C, :P], a[:, :, :C].transpose(1, 2))
    xs = rnd(torch.randn(B, S, cs, generator=g))
    gate = torch.rand(B, cs, generator=g)
    xd = xs.to(tdt).cuda()
    assert L.occd_channel_scale(xd.data_ptr(), gate.cuda().data_ptr(), code, B, S, cs, cs, st) == 0
    want = xs * gate[:, None, :]
    assert float((xd.float().cpu() - want).abs().max()) <= tol * float(want.abs().max())


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

