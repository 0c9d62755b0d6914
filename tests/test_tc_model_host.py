"""CPU check of the per-tap conv plans (csrc/conv.cu tc_geometry) through a host model of conv_tc_kernel
(tests/host_emul/tc_model.cu): tiling, TMA box origins / strides / zero fill, pipeline items, weight row addressing,
lane <-> position mapping -- for the default kernel (validated on the GPU) and its x-packed variant (TCX)."""
import ctypes as C
import os
import shutil
import subprocess

import pytest
import torch
import torch.nn.functional as F

from occdepth_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def model(tmp_path_factory):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    out = tmp_path_factory.mktemp("tc_model") / "libtc_model.so"
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O1", "-std=c++17", "--expt-relaxed-constexpr",
           "-shared", "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "occdepth_b200", "csrc"),
           "-o", str(out), os.path.join(ROOT, "tests", "host_emul", "tc_model.cu")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    lib = C.CDLL(str(out))
    lib.tc_model.restype = C.c_int
    lib.tc_model.argtypes = [C.POINTER(_lib.ConvDesc), C.c_int, C.c_void_p, C.POINTER(C.c_int)]
    lib.tc_model_error.restype = C.c_char_p
    return lib


def run_model(lib, xs, w, bias, stride, pad, mode, per_image=False):
    """xs: list of [B,Ci,D,H,W] sources (torch.cat along channels is what the conv sees); w [Cout, sum Ci, kd,kh,kw]
    (or [B, Cout, ...] when per_image)"""
    B, _, D, H, W = xs[0].shape
    wl = w if per_image else w[None]
    Cout = wl.shape[1]
    kd, kh, kw = wl.shape[3:]
    Cout_pad = (Cout + 15) // 16 * 16
    maxC = max(x.shape[1] for x in xs)
    KC = 64 if maxC > 32 else (32 if maxC > 16 else 16)
    Kpad = (maxC + KC - 1) // KC * KC
    d = _lib.ConvDesc()
    d.impl = {0: _lib.CONV_IMPL_TC, 1: _lib.CONV_IMPL_TCX, 2: _lib.CONV_IMPL_TCM2}[int(mode)]
    d.n_src = len(xs)
    keep = []
    for i, x in enumerate(xs):
        Ci = x.shape[1]
        cs = (Ci + 7) // 8 * 8
        xin = torch.zeros(B, D, H, W, cs)
        xin[..., :Ci] = x.permute(0, 2, 3, 4, 1)
        keep.append(xin)
        d.src[i] = xin.data_ptr()
        d.src_C[i], d.src_cstride[i], d.src_coff[i] = Ci, cs, 0
    d.B, d.ID, d.IH, d.IW = B, D, H, W
    OD, OH, OW = [(n + 2 * p - k) // s + 1 for n, p, k, s in zip((D, H, W), pad, (kd, kh, kw), stride)]
    for i in range(3):
        d.stride[i], d.omul[i], d.oadd[i] = stride[i], 1, 0
    taps, c0 = [], 0
    for si, x in enumerate(xs):
        Ci = x.shape[1]
        for a in range(kd):
            for b in range(kh):
                for c in range(kw):
                    taps.append((si, a - pad[0], b - pad[1], c - pad[2], wl[:, :, c0:c0 + Ci, a, b, c]))
        c0 += Ci
    nset = wl.shape[0]
    wp = torch.zeros(nset, len(taps), Cout_pad, Kpad)
    for i, t in enumerate(taps):
        wp[:, i, :Cout, :t[4].shape[2]] = t[4]
    bp = torch.zeros(Cout_pad)
    bp[:Cout] = bias
    d.n_taps = len(taps)
    for i, t in enumerate(taps):
        d.taps[i].src, d.taps[i].dz, d.taps[i].dy, d.taps[i].dx = t[0], t[1], t[2], t[3]
    d.weight, d.bias = wp.data_ptr(), bp.data_ptr()
    d.Cout, d.Cout_pad, d.Kpad = Cout, Cout_pad, Kpad
    d.weight_per_image = 1 if per_image else 0
    d.OD, d.OH, d.OW = OD, OH, OW
    d.ODf, d.OHf, d.OWf = OD, OH, OW
    out = torch.full((B, OD, OH, OW, Cout), float("nan"))
    d.out0 = out.data_ptr()
    d.out0_cstride = (Cout + 7) // 8 * 8
    info = (C.c_int * 8)()
    rc = lib.tc_model(C.byref(d), int(mode), out.data_ptr(), info)
    if rc != 0:
        return None, (rc, lib.tc_model_error().decode())
    keys = ("TD", "TH", "TW", "N_tile", "items", "stages", "group", "grid")
    return out.permute(0, 4, 1, 2, 3), dict(zip(keys, list(info)))


def reference(xs, w, bias, stride, pad, per_image):
    x = torch.cat(xs, 1)
    if not per_image:
        return F.conv3d(x, w, bias, stride, pad)
    return torch.cat([F.conv3d(x[b:b + 1], w[b], bias, stride, pad) for b in range(x.shape[0])])


CASES = {
    # name: (B, [Cin...], Cout, (D,H,W), k, stride, pad, per_image, xp_ok)
    "up1_like": (2, [40, 3], 48, (1, 13, 70), (1, 3, 3), (1, 1, 1), (0, 1, 1), False, True),
    "3d_c16": (1, [16], 16, (5, 7, 33), (3, 3, 3), (1, 1, 1), (1, 1, 1), False, True),
    "1d_w": (1, [64], 80, (3, 4, 61), (1, 1, 3), (1, 1, 1), (0, 0, 1), False, True),
    "stride_hd": (1, [24], 32, (6, 9, 31), (3, 3, 3), (2, 2, 1), (1, 1, 1), False, True),
    "stride2": (1, [20], 40, (6, 10, 14), (3, 3, 3), (2, 2, 2), (1, 1, 1), False, False),
    "pointwise_ntiles": (2, [72], 320, (1, 5, 9), (1, 1, 1), (1, 1, 1), (0, 0, 0), False, False),
    "per_image_1x1": (2, [96], 48, (1, 6, 11), (1, 1, 1), (1, 1, 1), (0, 0, 0), True, False),
    "per_image_3x3": (2, [32], 16, (1, 6, 35), (1, 3, 3), (1, 1, 1), (0, 1, 1), True, True),
}


@pytest.mark.parametrize("case", sorted(CASES))
def test_tc_m2_plan_model_matches_conv3d(model, case):
    """M2 plans (two M tiles per weight tile): pair walk incl. the phantom tile of an odd tile count, smem / TMEM
    budget of the doubled A stage"""
    B, Cins, Cout, dims, k, stride, pad, per_image, _ = CASES[case]
    g = torch.Generator().manual_seed(sum(Cins) + Cout + dims[2] + 7)
    xs = [torch.randn(B, c, *dims, generator=g) for c in Cins]
    wshape = (Cout, sum(Cins), *k)
    w = torch.randn(*((B,) + wshape if per_image else wshape), generator=g) / (sum(Cins) * k[0] * k[1] * k[2]) ** 0.5
    bias = torch.randn(Cout, generator=g)
    got, info = run_model(model, xs, w, bias, stride, pad, 2, per_image)
    if per_image:
        assert got is None and "m2" in info[1], info
        return
    assert got is not None, info
    ref = reference(xs, w, bias, stride, pad, per_image)
    assert torch.isfinite(got).all()
    assert torch.allclose(got, ref, rtol=1e-4, atol=1e-4), float((got - ref).abs().max())


@pytest.mark.parametrize("xp", [False, True])
@pytest.mark.parametrize("case", sorted(CASES))
def test_tc_plan_model_matches_conv3d(model, case, xp):
    B, Cins, Cout, dims, k, stride, pad, per_image, xp_ok = CASES[case]
    g = torch.Generator().manual_seed(sum(Cins) + Cout + dims[2])
    xs = [torch.randn(B, c, *dims, generator=g) for c in Cins]
    wshape = (Cout, sum(Cins), *k)
    w = torch.randn(*((B,) + wshape if per_image else wshape), generator=g) / (sum(Cins) * k[0] * k[1] * k[2]) ** 0.5
    bias = torch.randn(Cout, generator=g)
    got, info = run_model(model, xs, w, bias, stride, pad, xp, per_image)
    if xp and not xp_ok:
        assert got is None and "tcx" in info[1], info     # declined with a reason; auto mode keeps the per-tap kernel
        return
    assert got is not None, info
    ref = reference(xs, w, bias, stride, pad, per_image)
    assert got.shape == ref.shape
    assert torch.isfinite(got).all()
    assert torch.allclose(got, ref, rtol=1e-4, atol=1e-4), float((got - ref).abs().max())
    if xp:
        assert info["TW"] == 32 and info["N_tile"] == 3 * ((Cout + 15) // 16 * 16)
