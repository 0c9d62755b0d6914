"""SFA lift: CUDA kernel (through the drop-in module / C ABI) vs the CPU oracle on the same seeded inputs."""
import pytest
import torch

from oracle import functional as OF
from oracle import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dataset,V,P,C", [("kitti", 2, 1, 64), ("kitti", 1, 1, 32), ("NYU", 2, 1, 200),
                                           ("kitti", 2, 5, 16), ("kitti", 3, 1, 64), ("NYU", 1, 9, 36)])
def test_sfa_module_vs_oracle(dataset, V, P, C):
    from occdepth_b200.models.SFA import SFA
    torch.manual_seed(0)
    scene, ps = (24, 20, 12), 2
    S = [s // ps for s in scene]
    N = S[0] * S[1] * S[2]
    h, w = 23, 41
    x2d = torch.randn(V, C, h, w)
    pix, fov = synth.random_indices(N, w, h, n_views=V, P=P, seed=3, margin=(6, 5))
    want = OF.sfa(x2d, pix, fov, scene, dataset, ps)
    got = SFA(scene, dataset, ps)(x2d.cuda(), pix.cuda(), fov.cuda()).cpu()
    assert got.shape == want.shape
    # fp32 gather / multiply / rsqrt: tolerance 1e-5 absolute on O(1) features
    assert float((got - want).abs().max()) <= 1e-5 * max(1.0, float(want.abs().max()))


def test_sfa_empty_and_all_masked():
    from occdepth_b200.models.SFA import SFA
    scene, ps = (8, 8, 4), 1
    N = 8 * 8 * 4
    x2d = torch.randn(2, 8, 5, 7)
    pix = torch.zeros(2, N, 1, 2, dtype=torch.int64)
    fov = torch.zeros(2, N, 1, dtype=torch.bool)
    got = SFA(scene, "kitti", ps)(x2d.cuda(), pix.cuda(), fov.cuda())
    assert float(got.abs().max()) == 0.0


@pytest.mark.parametrize("precision", ["tf32", "bf16"])
def test_lift_multiscale_kitti_calibration(precision):
    """fused 4-scale / 2-view production path (features and output in the plan's element type) against the oracle
    run on the same rounded feature maps, with the KITTI-like vox2pix indices (realistic locality, FOV fraction ~0.7)."""
    import gpu_cases as G
    from occdepth_b200.engine import CL
    from occdepth_b200.models.SFA import lift_multiscale
    H, W, C = 94, 343, 64
    full, ps = (64, 64, 8), 2
    pix, fov, _, _ = synth.kitti_indices(W, H, full, ps, voxel=0.8)
    g = torch.Generator().manual_seed(0)
    feats, x_rgb = [], [{}, {}]
    for s in (1, 2, 4, 8):
        h, w = synth.feature_hw(H, W, s)
        f = G.rnd(precision)(torch.randn(2, C, h, w, generator=g))
        feats.append(f.permute(0, 2, 3, 1).contiguous().to(torch.float32 if precision == "tf32" else torch.bfloat16).cuda())
        for v in range(2):
            x_rgb[v]["1_%d" % s] = f[v].float()
    want = OF.lift_flosp(x_rgb, pix, fov, ["1", "2", "4", "8"], full, "kitti", ps)
    X, Y, Z = [s // ps for s in full]
    out = CL.alloc(1, X, Y, Z, C, torch.device("cuda"), precision=precision)
    lift_multiscale(feats, [1, 2, 4, 8], pix.cuda(), fov.cuda(), out, "kitti", full, ps)
    got = out.to_planar()[0].cpu()
    assert fov.float().mean() > 0.3
    # the output is rounded once to the plan's element type
    assert float((got - want).abs().max()) <= G.TOL[precision] * float(want.abs().max())
