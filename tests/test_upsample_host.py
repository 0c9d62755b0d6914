"""CPU check of the rows-decomposed bilinear resize (csrc/upsample_rows.cuh) through its host emulation vs
F.interpolate(mode="bilinear", align_corners=True) (UpSampleBN.forward, unet2d.py:39-44)."""
import ctypes as C
import os
import shutil
import subprocess

import pytest
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    out = tmp_path_factory.mktemp("up_emul") / "libup_emul.so"
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O1", "-std=c++17", "--expt-relaxed-constexpr",
           "-shared", "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "occdepth_b200", "csrc"),
           "-o", str(out), os.path.join(ROOT, "tests", "host_emul", "upsample_emul.cu")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    lib = C.CDLL(str(out))
    lib.upsample_rows_emulate.restype = C.c_int
    lib.upsample_rows_emulate.argtypes = [C.c_void_p] * 2 + [C.c_int] * 10
    return lib


@pytest.mark.parametrize("B,Cn,h,w,OH,OW,in_off,out_off", [
    (1, 16, 5, 7, 10, 14, 0, 0),
    (2, 160, 6, 9, 12, 17, 0, 0),          # 20 vectors -> CVB 32 with idle lanes
    (1, 24, 7, 5, 13, 11, 8, 16),          # channel windows inside wider buffers, odd target sizes
    (1, 8, 1, 1, 4, 6, 0, 0),              # single source pixel
    (1, 320, 3, 4, 3, 4, 0, 8),            # identity size, two channel blocks
])
def test_rows_upsample_matches_interpolate(emul, B, Cn, h, w, OH, OW, in_off, out_off):
    g = torch.Generator().manual_seed(Cn + h + OW)
    cs_in, cs_out = in_off + Cn + 8, out_off + Cn
    x = torch.randn(B, h, w, cs_in, generator=g).to(torch.bfloat16)
    out = torch.full((B, OH, OW, cs_out), float("nan")).to(torch.bfloat16)
    rc = emul.upsample_rows_emulate(x.data_ptr(), out.data_ptr(), B, h, w, OH, OW, Cn, cs_in, in_off, cs_out, out_off)
    assert rc == 0
    src = x[..., in_off:in_off + Cn].float().permute(0, 3, 1, 2)
    ref = F.interpolate(src, size=(OH, OW), mode="bilinear", align_corners=True).permute(0, 2, 3, 1)
    got = out[..., out_off:out_off + Cn].float()
    assert torch.isfinite(got).all()
    assert torch.allclose(got, ref, rtol=2 ** -7, atol=2 ** -7), float((got - ref).abs().max())
    if out_off:
        assert torch.isnan(out[..., :out_off].float()).all()         # channels outside the window are untouched
