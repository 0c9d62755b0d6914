"""CPU check of the strip-decomposed SE gate fold (csrc/se_fold_strip.cuh) through its host emulation:
gate = sigmoid(W2 hidden + b2) folded into the per-image projection weights (geffnet SqueezeExcite + conv_pwl,
unet2d.py:188-196)."""
import ctypes as C
import os
import shutil
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    out = tmp_path_factory.mktemp("se_emul") / "libse_emul.so"
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O1", "-std=c++17", "--expt-relaxed-constexpr",
           "-shared", "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "occdepth_b200", "csrc"),
           "-o", str(out), os.path.join(ROOT, "tests", "host_emul", "se_fold_emul.cu")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    lib = C.CDLL(str(out))
    lib.se_fold_strip_emulate.restype = C.c_int
    lib.se_fold_strip_emulate.argtypes = [C.c_void_p] * 6 + [C.c_int] * 5
    return lib


@pytest.mark.parametrize("B,Cn,R,rows,Kpad", [(1, 32, 8, 32, 64), (2, 192, 12, 48, 192), (2, 288, 12, 48, 320),
                                              (1, 1344, 56, 224, 1344), (2, 40, 3, 16, 64)])
def test_strip_fold_matches_torch(emul, B, Cn, R, rows, Kpad):
    g = torch.Generator().manual_seed(Cn + R)
    hidden = torch.randn(B, R, generator=g)
    w2 = torch.randn(Cn, R, generator=g) / R ** 0.5
    b2 = torch.randn(Cn, generator=g)
    master = torch.zeros(rows, Kpad)
    master[:, :Cn] = torch.randn(rows, Cn, generator=g)
    pool = torch.full((B, Cn), 12345, dtype=torch.int64)
    out = torch.full((B, rows, Kpad), float("nan")).to(torch.bfloat16)
    w2t = w2.t().contiguous()
    rc = emul.se_fold_strip_emulate(pool.data_ptr(), hidden.data_ptr(), w2t.data_ptr(), b2.data_ptr(),
                                    master.data_ptr(), out.data_ptr(), B, Cn, R, rows, Kpad)
    assert rc == 0
    gate = torch.sigmoid(hidden @ w2.t() + b2)                       # [B][C]
    want = torch.zeros(B, rows, Kpad)
    want[:, :, :Cn] = master[None, :, :Cn] * gate[:, None, :]
    got = out.float()
    assert torch.isfinite(got).all()                                 # pad columns are written (as zeros)
    assert torch.allclose(got, want.to(torch.bfloat16).float(), rtol=2 ** -7, atol=1e-6)
    assert (got[:, :, Cn:] == 0).all()
    assert (pool == 0).all()                                         # squeeze sums cleared for the next forward
