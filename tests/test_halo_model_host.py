"""CPU check of the halo-tile conv plans (csrc/conv.cu halo_geometry) through a host model of conv_halo_kernel
(tests/host_emul/halo_model.cu): the real tile / box / tap-offset / lane-mapping arithmetic of the plan is run on
the CPU and compared with torch's conv3d -- for the 27-tap kernel that is validated on the GPU and for the
x-packed variant (three W taps per MMA, lane-shifted epilogue sum)."""
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
    out = tmp_path_factory.mktemp("halo_model") / "libhalo_model.so"
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O1", "-std=c++17", "--expt-relaxed-constexpr",
           "-shared", "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "occdepth_b200", "csrc"),
           "-o", str(out), os.path.join(ROOT, "tests", "host_emul", "halo_model.cu")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    lib = C.CDLL(str(out))
    lib.halo_model.restype = C.c_int
    lib.halo_model.argtypes = [C.POINTER(_lib.ConvDesc), C.c_int, C.c_void_p, C.POINTER(C.c_int)]
    lib.halo_model_error.restype = C.c_char_p
    return lib


def run_model(lib, x, w, bias, dil, xp, ksize=(3, 3, 3)):
    """x [B,Cin,D,H,W] fp32, w [Cout,Cin,kd,kh,kw]; returns (out [B,Cout,D,H,W], info) or (None, error)"""
    B, Cin, D, H, W = x.shape
    Cout = w.shape[0]
    Cout_pad = (Cout + 15) // 16 * 16
    KC = 64 if Cin > 32 else (32 if Cin > 16 else 16)
    cs = (Cin + 7) // 8 * 8
    xin = torch.zeros(B, D, H, W, cs)
    xin[..., :Cin] = x.permute(0, 2, 3, 4, 1)
    taps = []
    kd, kh, kw = ksize
    for a in range(kd):
        for b in range(kh):
            for c in range(kw):
                taps.append(((a - kd // 2) * dil, (b - kh // 2) * dil, (c - kw // 2) * dil, w[:, :, a, b, c]))
    wp = torch.zeros(len(taps), Cout_pad, KC)
    for i, t in enumerate(taps):
        wp[i, :Cout, :Cin] = t[3]
    bp = torch.zeros(Cout_pad)
    bp[:Cout] = bias
    d = _lib.ConvDesc()
    d.impl = _lib.CONV_IMPL_HALOX if xp else _lib.CONV_IMPL_HALO
    d.n_src = 1
    d.src[0] = xin.data_ptr()
    d.src_C[0], d.src_cstride[0], d.src_coff[0] = Cin, cs, 0
    d.B, d.ID, d.IH, d.IW = B, D, H, W
    for i in range(3):
        d.stride[i], d.omul[i], d.oadd[i] = 1, 1, 0
    d.n_taps = len(taps)
    for i, t in enumerate(taps):
        d.taps[i].src, d.taps[i].dz, d.taps[i].dy, d.taps[i].dx = 0, t[0], t[1], t[2]
    d.weight, d.bias = wp.data_ptr(), bp.data_ptr()
    d.Cout, d.Cout_pad, d.Kpad = Cout, Cout_pad, KC
    d.OD, d.OH, d.OW = D, H, W
    d.ODf, d.OHf, d.OWf = D, H, W
    out = torch.full((B, D, H, W, Cout), float("nan"))
    d.out0 = out.data_ptr()       # only checked for non-null by the geometry code
    d.out0_cstride = (Cout + 7) // 8 * 8
    info = (C.c_int * 10)()
    rc = lib.halo_model(C.byref(d), 1 if xp else 0, out.data_ptr(), info)
    if rc != 0:
        return None, (rc, lib.halo_model_error().decode())
    keys = ("BD", "BH", "BW", "PD", "PH", "PW", "nM", "N_tile", "groups", "stages")
    return out.permute(0, 4, 1, 2, 3), dict(zip(keys, list(info)))


CASES = [
    # B, Cin, Cout, D,  H,  W, dil, ksize
    (1, 32, 32, 5, 9, 40, 1, (3, 3, 3)),
    (2, 32, 20, 4, 11, 33, 1, (3, 3, 3)),     # Cout 20 -> pad 32 (conv_classes), ragged last W tile
    (1, 32, 32, 7, 13, 64, 2, (3, 3, 3)),     # dilation 2: 8 residue sub-grids
    (1, 32, 32, 7, 10, 35, 3, (3, 3, 3)),     # dilation 3, extents not multiples of 3
    (1, 32, 2, 3, 8, 31, 1, (3, 3, 3)),       # occ_classes: Cout 2 -> pad 16, N = 48
    (1, 16, 16, 1, 12, 45, 1, (1, 3, 3)),     # 2-D
    (1, 64, 48, 1, 9, 38, 1, (1, 1, 3)),      # 1-D along W, KC = 64
    (1, 8, 32, 6, 6, 6, 1, (3, 3, 3)),        # tiny grid
]


@pytest.mark.parametrize("xp", [False, True])
@pytest.mark.parametrize("B,Cin,Cout,D,H,W,dil,ksize", CASES)
def test_halo_plan_model_matches_conv3d(model, B, Cin, Cout, D, H, W, dil, ksize, xp):
    g = torch.Generator().manual_seed(Cin * 100 + Cout + W + dil)
    x = torch.randn(B, Cin, D, H, W, generator=g)
    w = torch.randn(Cout, Cin, *ksize, generator=g) / (Cin * ksize[0] * ksize[1] * ksize[2]) ** 0.5
    bias = torch.randn(Cout, generator=g)
    got, info = run_model(model, x, w, bias, dil, xp, ksize)
    if got is None:
        # the plain halo scheme may decline a shape (< 50 % useful rows ...): auto mode then uses the per-tap kernel
        assert not xp or "x-packed" not in info[1], info
        pytest.skip("plan declined: %s" % (info,))
    pad = tuple(dil * (k // 2) for k in ksize)
    ref = F.conv3d(x, w, bias, padding=pad, dilation=dil)
    assert torch.isfinite(got).all(), "a stored output depends on rows outside the loaded box"
    assert torch.allclose(got, ref, rtol=1e-4, atol=1e-4), float((got - ref).abs().max())
    if xp:
        assert info["PW"] == 32 and info["BW"] == 30 and info["N_tile"] == 3 * ((Cout + 15) // 16 * 16)
        assert info["groups"] == ksize[0] * ksize[1]
