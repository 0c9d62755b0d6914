"""Pins oracle/functional.py against the UNMODIFIED reference modules (CPU fp32, seeded weights).

Runs only where /root/reference exists (the build container); elsewhere the committed fixtures in
tests/golden/ (generated from the same reference runs, oracle/gen_golden.py) carry the pin.
"""
import pytest
import torch
import torch.nn as nn

from oracle import functional as OF
from oracle import ref_import, synth

pytestmark = pytest.mark.reference
TOL = dict(rtol=1e-4, atol=1e-5)


@pytest.fixture(scope="module")
def ref():
    return ref_import.modules()


def _close(a, b, **kw):
    tol = dict(TOL)
    tol.update(kw)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert torch.allclose(a, b, **tol), float((a - b).abs().max())


@pytest.mark.parametrize("dataset,V,P", [("kitti", 2, 1), ("kitti", 1, 1), ("NYU", 2, 1), ("kitti", 2, 5), ("kitti", 3, 1)])
def test_sfa(ref, dataset, V, P):
    torch.manual_seed(0)
    scene = (16, 12, 8)
    ps = 2
    S = [s // ps for s in scene]
    N = S[0] * S[1] * S[2]
    C, h, w = 16, 11, 23
    x2d = torch.randn(V, C, h, w)
    pix, fov = synth.random_indices(N, w, h, n_views=V, P=P, seed=3, margin=(5, 4))
    m = ref.SFA.SFA(scene, dataset, ps)
    want = m(x2d, pix.clone(), fov.clone())
    got = OF.sfa(x2d, pix, fov, scene, dataset, ps)
    _close(got, want.contiguous(), rtol=1e-5, atol=1e-6)


def test_bottleneck_process_downsample(ref):
    torch.manual_seed(0)
    with ref_import.quiet():
        proc = ref.modules.Process(16, nn.BatchNorm3d, 0.1, dilations=[1, 2, 3]).eval()
        down = ref.modules.Downsample(16, nn.BatchNorm3d, 0.1).eval()
        up = ref.modules.Upsample(32, 16, nn.BatchNorm3d, 0.1).eval()
    for m in (proc, down, up):
        synth.randomize_bn_(m)
    x = torch.randn(1, 16, 8, 10, 6)
    with torch.no_grad():
        _close(OF.process({"p." + k: v for k, v in proc.state_dict().items()}, "p", x), proc(x))
        y = down(x)
        _close(OF.downsample({"p." + k: v for k, v in down.state_dict().items()}, "p", x), y)
        _close(OF.upsample({"p." + k: v for k, v in up.state_dict().items()}, "p", y), up(y))


@pytest.mark.parametrize("which", ["kitti", "nyu"])
def test_unet3d(ref, which):
    torch.manual_seed(0)
    with ref_import.quiet():
        if which == "kitti":
            full, ps, f = (32, 32, 16), 2, 16
            m = ref.unet3d_kitti.UNet3D(5, nn.BatchNorm3d, full, f, ps, context_prior=True, cascade_cls=True,
                                        occluded_cls=True).eval()
            x = torch.randn(1, f, 16, 16, 8)
        else:
            full, f = (12, 8, 12), 16
            m = ref.unet3d_nyu.UNet3D(5, nn.BatchNorm3d, f, full, context_prior=True, cascade_cls=False).eval()
            x = torch.randn(1, f, 12, 8, 12)
    synth.randomize_bn_(m)
    sd = {"n." + k: v for k, v in m.state_dict().items()}
    with torch.no_grad():
        want = m({"x3d": x})
        if which == "kitti":
            got = OF.unet3d_kitti(sd, "n", x, full, ps, True, True, True)
        else:
            got = OF.unet3d_nyu(sd, "n", x, full, 4, True, False)
    assert set(got.keys()) == set(want.keys())
    for k in want:
        _close(got[k], want[k], rtol=1e-3, atol=1e-4)


def test_unet2d(ref):
    torch.manual_seed(0)
    with ref_import.quiet():
        m = ref.unet2d.UNet2D.build(out_feature=16, use_decoder=True, backbone_2d_name="tf_efficientnet_b3_ns",
                                    return_up_feats=1).eval()
    synth.randomize_bn_(m)
    sd = {"net_rgb." + k: v for k, v in m.state_dict().items()}
    x = torch.randn(1, 3, 38, 45)
    with torch.no_grad():
        want = m(x)
        got = OF.unet2d(sd, "net_rgb", x, "tf_efficientnet_b3_ns", 1)
    assert set(got.keys()) == set(want.keys())
    for k in want:
        _close(got[k], want[k], rtol=1e-3, atol=1e-4)


def test_occdepth_forward_small(ref):
    """whole OccDepth.forward ("flosp", kitti decoder with CRP + cascade) on a tiny stereo pair."""
    torch.manual_seed(0)
    full = (32, 32, 16)
    cfg = synth.occdepth_cfg(full_scene_size=full, feature=16, feature_2d_oc=16, n_classes=6,
                             backbone_2d_name="tf_efficientnet_b3_ns")
    with ref_import.quiet():
        m = ref.OccDepth.OccDepth(class_names=["c"] * 6, class_weights=torch.ones(6), full_scene_size=full,
                                  project_res=["1", "2", "4", "8"], config=cfg).eval()
    synth.randomize_bn_(m)
    H, W = 33, 49
    g = torch.Generator().manual_seed(0)
    img = torch.randn(1, 2, 3, H, W, generator=g)
    N = 16 * 16 * 8
    pix, fov = synth.random_indices(N, W, H, n_views=2, P=1, seed=5, margin=(10, 6))
    batch = {"img": img, "projected_pix_2": [pix], "fov_mask_2": [fov]}
    with torch.no_grad():
        want = m(batch)
        ocfg = dict(cfg)
        ocfg["project_res"] = ["1", "2", "4", "8"]
        got = OF.occdepth_forward(m.state_dict(), batch, ocfg)
    assert set(got.keys()) == set(want.keys())
    for k in want:
        _close(got[k], want[k], rtol=2e-3, atol=2e-4)


def test_flosp_depth(ref):
    """FlospDepth (DepthNet + frustum grid + trilinear sampling + masked mean), kitti geometry, 2 cameras"""
    import copy
    import occdepth.models.flosp_depth as rfd
    torch.manual_seed(0)
    conf = copy.deepcopy(rfd.flosp_depth_conf_map["kitti"])
    H, W = 96, 320
    conf.update(scene_size=(64, 64, 8), project_scale=2, return_depth=True, final_dim=(H, W), output_channels=16,
                x_bound=[0, 12.8, 0.2], y_bound=[-6.4, 6.4, 0.2], z_bound=[-2, -0.4, 0.2], d_bound=[2.0, 18.0, 0.5],
                depth_net_conf=dict(in_channels=16, mid_channels=32))
    with ref_import.quiet():
        m = ref.flosp_depth.FlospDepth(**conf).eval()
    synth.seed_weights_(m, 5)
    K, Ts = synth.kitti_calib(W, H, focal=220.0)
    cam_k = [torch.from_numpy(K).unsqueeze(0).repeat(2, 1, 1)]
    T = [torch.stack([torch.from_numpy(t) for t in Ts])]
    ida = [torch.eye(4).unsqueeze(0).repeat(2, 1, 1)]
    feat = torch.randn(1, 2, 16, H // 8, W // 8)
    with torch.no_grad():
        want, want_d = m(feat, cam_k, T, ida, None)
        got, got_d = OF.flosp_depth({"f." + k: v for k, v in m.state_dict().items()}, "f", feat, cam_k, T, ida, conf)
    assert float((want > 0).float().mean()) > 0.2     # the frustum actually covers part of the grid
    _close(got_d, want_d, rtol=1e-4, atol=1e-6)
    _close(got, want, rtol=1e-3, atol=1e-5)


def _nyu_virtual_batch(H, W, full, seed=0):
    g = torch.Generator().manual_seed(seed)
    img = torch.randn(1, 1, 3, H, W, generator=g)
    depth = torch.rand(1, 1, H, W, generator=g) * 7.5 + 0.5
    depth[0, 0, 3, 5] = 0.0          # exercises the inf -> 0 disparity rule (OccDepth.py:248)
    N = full[0] * full[1] * full[2]
    pix, fov = synth.random_indices(N, W, H, n_views=2, P=1, seed=seed + 1, margin=(8, 6))
    return {"img": img, "gt_depth": depth, "virtual_bf": [torch.tensor(51.88579 * W / 640.0)],
            "vox_origin": torch.zeros(1, 3, dtype=torch.float64), "projected_pix_1": [pix], "fov_mask_1": [fov]}


def test_occdepth_forward_nyu_virtual_view(ref):
    """NYU-type path: one real view, the right view synthesised from gt_depth (generate_virtual_img)"""
    torch.manual_seed(0)
    full = (12, 8, 12)
    cfg = synth.occdepth_cfg(dataset="NYU", full_scene_size=full, project_scale=1, feature=16, feature_2d_oc=16,
                             n_classes=6, cascade_cls=False, backbone_2d_name="tf_efficientnet_b3_ns")
    with ref_import.quiet():
        m = ref.OccDepth.OccDepth(["c"] * 6, torch.ones(6), full_scene_size=full, project_res=["1", "2", "4", "8"],
                                  config=cfg).eval()
    synth.seed_weights_(m, 3)
    batch = _nyu_virtual_batch(32, 64, full)
    ocfg = dict(cfg)
    ocfg["project_res"] = ["1", "2", "4", "8"]
    with torch.no_grad():
        want = m(batch)
        got = OF.occdepth_forward(m.state_dict(), batch, ocfg)
    assert set(got.keys()) == set(want.keys())
    for k in want:
        _close(got[k], want[k], rtol=2e-3, atol=2e-4)


@pytest.mark.parametrize("pattern_id", [0, 1, 3, 6, 8])
def test_vox2pix_against_the_numba_reference(pattern_id):
    """oracle/projection.py vs occdepth.data.utils.helpers.vox2pix itself (numba-compiled), bit for bit"""
    import numpy as np
    from oracle import gen_golden, projection
    helpers = ref_import.data_helpers()
    for name, E, K, org, vs, W, H, scene, _ in gen_golden.vox2pix_cases():
        want = helpers.vox2pix(E, K, org, vs, W, H, scene, pattern_id)
        got = projection.vox2pix(E, K, org, vs, W, H, scene, pattern_id)
        for g, w in zip(got, want):
            assert g.dtype == w.dtype and np.array_equal(g, w, equal_nan=(g.dtype.kind == "f")), name
