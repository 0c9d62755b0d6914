"""Implicit-GEMM convolution through the C ABI: tcgen05 path and SIMT path vs torch fp32 on bf16-rounded operands."""
import pytest

import gpu_cases as G

pytestmark = pytest.mark.gpu
# operands are exactly representable in bf16, accumulation is fp32; the only rounding is the bf16 store
TOL = 2 ** -7


@pytest.mark.parametrize("impl", ["tc", "simt"])
@pytest.mark.parametrize("case", sorted(G.CONV_CASES))
def test_conv(impl, case):
    from occdepth_b200 import _lib
    e, info = G.conv_case(_lib.CONV_IMPL_TC if impl == "tc" else _lib.CONV_IMPL_SIMT, **G.CONV_CASES[case])
    assert e <= TOL, (e, info)


@pytest.mark.parametrize("impl", ["tc", "simt"])
def test_conv_transpose_and_multi(impl):
    from occdepth_b200 import _lib
    i = _lib.CONV_IMPL_TC if impl == "tc" else _lib.CONV_IMPL_SIMT
    e, info = G.convT_case(i)
    assert e <= TOL, (e, info)
    e, info = G.multi_case(i)
    assert e <= TOL, (e, info)


def test_conv_large_vs_simt():
    """full-size head conv shape (Cin=Cout=32, dil 3) on a 64x64x32 slab: TC vs SIMT on the same buffers."""
    import torch
    from occdepth_b200 import _lib
    from occdepth_b200.engine import CL, Plan
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(1)
    x = CL.from_planar(torch.randn(1, 32, 64, 64, 32, generator=g).cuda())
    w = (torch.randn(32, 32, 3, 3, 3, generator=g) / 30).cuda()
    b = torch.randn(32, generator=g).cuda()
    outs = []
    for impl in (_lib.CONV_IMPL_TC, _lib.CONV_IMPL_SIMT):
        plan = Plan(dev)
        outs.append(plan.conv(x, w, b, padding=3, dilation=3, act="relu", impl=impl))
        plan.run()
    torch.cuda.synchronize()
    a, c = outs[0].buf.float(), outs[1].buf.float()
    assert float((a - c).abs().max()) <= TOL * float(c.abs().max())


@pytest.mark.parametrize("case", sorted(G.HALO_CASES))
def test_conv_halo(case):
    """halo-tile tcgen05 kernel (row-shifted swizzled smem views, sub-sampled TMA for dilation)"""
    from occdepth_b200 import _lib
    e, info = G.conv_case(_lib.CONV_IMPL_HALO, **G.HALO_CASES[case])
    assert e <= TOL, (e, info)


@pytest.mark.parametrize("case", sorted(G.HALO_CASES))
def test_conv_halo_xpacked(case):
    """x-packed halo kernel (three W taps per MMA, lane-shifted epilogue).  Plan geometry and epilogue mapping are
    covered on the CPU (tests/test_halo_model_host.py); the first GPU run is opt-in until it has been seen green."""
    import os
    from occdepth_b200 import _lib
    if os.environ.get("OCCD_EXPERIMENTAL") != "1":
        pytest.skip("x-packed halo kernel: set OCCD_EXPERIMENTAL=1 to run")
    try:
        e, info = G.conv_case(_lib.CONV_IMPL_HALOX, **G.HALO_CASES[case])
    except RuntimeError as err:
        if "(halo)" in str(err):        # geometry declined (e.g. weights + stage do not fit): auto mode falls back
            pytest.skip(str(err))
        raise
    assert e <= TOL, (e, info)


TCX_CASES = {
    "x_2d_multi": dict(k=(1, 3, 3), Cin=40, Cout=48, dims=(1, 13, 70), B=2),
    "x_3d_c16": dict(Cin=16, Cout=16, dims=(5, 7, 33)),
    "x_1d_w_c64": dict(k=(1, 1, 3), Cin=64, Cout=80, dims=(3, 4, 61), act="leaky"),
    "x_stride_hd": dict(Cin=24, Cout=32, dims=(6, 9, 31), stride=(2, 2, 1), res=False),
    "x_big_2d": dict(k=(1, 3, 3), Cin=80, Cout=80, dims=(1, 90, 200), res=True),
}


@pytest.mark.parametrize("case", sorted(TCX_CASES))
def test_conv_tc_xpacked(case):
    """x-packed per-tap kernel (TCX); opt-in until it has been seen green on a B200 (CPU model:
    tests/test_tc_model_host.py)"""
    import os
    from occdepth_b200 import _lib
    if os.environ.get("OCCD_EXPERIMENTAL") != "1":
        pytest.skip("x-packed per-tap kernel: set OCCD_EXPERIMENTAL=1 to run")
    e, info = G.conv_case(_lib.CONV_IMPL_TCX, **TCX_CASES[case])
    assert e <= TOL, (e, info)


M2_CASES = {
    "m_2d_c160": dict(k=(1, 3, 3), Cin=160, Cout=160, dims=(1, 40, 90), res=True),       # N 160: one TMEM set
    "m_2d_c64_n128": dict(k=(1, 3, 3), Cin=64, Cout=128, dims=(1, 33, 70), B=2),        # N 128: two sets
    "m_3d_odd_tiles": dict(Cin=32, Cout=64, dims=(3, 9, 15)),                            # odd tile count: phantom tile
    "m_1x1_n256": dict(k=(1, 1, 1), Cin=96, Cout=256, dims=(1, 30, 50), act="silu"),
    "m_ntiles2": dict(k=(1, 3, 3), Cin=48, Cout=320, dims=(1, 20, 37)),                  # two N tiles of 160
}


@pytest.mark.parametrize("case", sorted(M2_CASES))
def test_conv_tc_m2(case):
    """M2 per-tap kernel (two M tiles per weight tile); opt-in until it has been seen green on a B200 (CPU model:
    tests/test_tc_model_host.py)"""
    import os
    from occdepth_b200 import _lib
    if os.environ.get("OCCD_EXPERIMENTAL") != "1":
        pytest.skip("M2 per-tap kernel: set OCCD_EXPERIMENTAL=1 to run")
    e, info = G.conv_case(_lib.CONV_IMPL_TCM2, **M2_CASES[case])
    assert e <= TOL, (e, info)
