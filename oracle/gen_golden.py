"""Generates tests/golden/*.pt by running the UNMODIFIED reference (imported from /root/reference through
oracle/ref_import.py) on seeded inputs and weights.  Run in the build container:  python -m oracle.gen_golden

Each fixture stores the inputs (or the seed that regenerates them), the seed of `synth.seed_weights_`, and the
reference outputs (fp16-safe small tensors kept as fp32).
"""
import os

import torch
import torch.nn as nn

from . import ref_import, synth

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def sfa_cases():
    return [("kitti", 2, 1, 16), ("NYU", 2, 1, 16), ("kitti", 1, 1, 8), ("kitti", 2, 5, 8), ("kitti", 3, 1, 8)]


def gen_sfa(ref):
    out = {}
    scene, ps, h, w = (12, 10, 8), 2, 9, 17
    N = 6 * 5 * 4
    for i, (ds, V, P, C) in enumerate(sfa_cases()):
        g = torch.Generator().manual_seed(100 + i)
        x2d = torch.randn(V, C, h, w, generator=g)
        pix, fov = synth.random_indices(N, w, h, n_views=V, P=P, seed=200 + i, margin=(4, 3))
        y = ref.SFA.SFA(scene, ds, ps)(x2d, pix.clone(), fov.clone()).contiguous()
        out["case%d" % i] = dict(dataset=ds, x2d=x2d, pix=pix, fov=fov, out=y, scene=scene, ps=ps)
    torch.save(out, os.path.join(OUT, "sfa.pt"))


def gen_unet3d(ref):
    with ref_import.quiet():
        mk = ref.unet3d_kitti.UNet3D(5, nn.BatchNorm3d, (32, 32, 16), 16, 2, context_prior=True, cascade_cls=True,
                                     occluded_cls=True).eval()
        mn = ref.unet3d_nyu.UNet3D(5, nn.BatchNorm3d, 16, (12, 8, 12), context_prior=True, cascade_cls=False).eval()
    synth.seed_weights_(mk, 11)
    synth.seed_weights_(mn, 12)
    g = torch.Generator().manual_seed(3)
    xk = torch.randn(1, 16, 16, 16, 8, generator=g)
    xn = torch.randn(1, 16, 12, 8, 12, generator=g)
    with torch.no_grad():
        ok, on = mk({"x3d": xk}), mn({"x3d": xn})
    keep = ("ssc_logit", "occ_logit", "occluded_logit", "P_logits")
    torch.save(dict(kitti=dict(x=xk, seed=11, out={k: v for k, v in ok.items() if k in keep}),
                    nyu=dict(x=xn, seed=12, out={k: v for k, v in on.items() if k in keep})),
               os.path.join(OUT, "unet3d.pt"))


def gen_occdepth(ref):
    full = (32, 32, 16)
    cfg = synth.occdepth_cfg(full_scene_size=full, feature=16, feature_2d_oc=16, n_classes=6,
                             backbone_2d_name="tf_efficientnet_b3_ns")
    with ref_import.quiet():
        m = ref.OccDepth.OccDepth(["c"] * 6, torch.ones(6), full_scene_size=full, project_res=["1", "2", "4", "8"],
                                  config=cfg).eval()
    synth.seed_weights_(m, 21)
    H, W = 33, 49
    g = torch.Generator().manual_seed(0)
    img = torch.randn(1, 2, 3, H, W, generator=g)
    pix, fov = synth.random_indices(16 * 16 * 8, W, H, n_views=2, P=1, seed=5, margin=(10, 6))
    with torch.no_grad():
        out = m({"img": img, "projected_pix_2": [pix], "fov_mask_2": [fov]})
        feats = m.net_rgb(img[:, 0])
    torch.save(dict(img=img, pix=pix, fov=fov, seed=21, cfg=dict(cfg), ssc_logit=out["ssc_logit"],
                    occ_logit=out["occ_logit"], feat_1_8=feats["1_8"], feat_1_1_mean=feats["1_1"].mean((2, 3))),
               os.path.join(OUT, "occdepth_small.pt"))


def vox2pix_cases():
    """(name, cam_E, cam_k, vox_origin, voxel_size, img_W, img_H, scene_size, pattern_id): a KITTI-like rig with a
    slightly rotated pose (so every term of the pose matrix matters), an NYU-like one, and a camera INSIDE the volume
    (voxels behind it and one column of centres in the camera plane: z <= 0, division by zero)"""
    import math
    import numpy as np

    def rot(ax, a):
        c, s = math.cos(a), math.sin(a)
        R = np.eye(3)
        i, j = [(1, 2), (0, 2), (0, 1)][ax]
        R[i, i] = c; R[j, j] = c; R[i, j] = -s; R[j, i] = s
        return R
    base = np.array([[0, -1, 0], [0, 0, -1], [1, 0, 0.0]])
    E1 = np.eye(4)
    E1[:3, :3] = rot(0, 0.03) @ rot(1, -0.02) @ rot(2, 0.01) @ base
    E1[:3, 3] = [0.011, -0.083, -0.271]
    K1 = np.array([[707.0912, 0, 601.8873], [0, 707.0912, 183.1104], [0, 0, 1.0]])
    E2 = np.eye(4)
    E2[:3, :3] = rot(0, -0.4) @ rot(2, 0.7)
    E2[:3, 3] = [0.3, -0.2, 1.1]
    K2 = np.array([[518.8579, 0, 320.0], [0, 518.8579, 240.0], [0, 0, 1.0]])
    E3 = np.eye(4)
    E3[:3, :3] = base
    E3[:3, 3] = [0.0, 0.0, -0.75]         # cam z = lidar x - 0.75: the centres at x = 0.75 lie exactly in the camera plane
    return [("kitti_p0", E1, K1, np.array([0, -6.4, -2.0]), 0.4, 1220, 370, (12.8, 12.8, 3.2), 0),
            ("kitti_p3", E1, K1, np.array([0, -6.4, -2.0]), 0.8, 1220, 370, (12.8, 12.8, 3.2), 3),
            ("nyu_p8", E2.astype(np.float32), K2, np.array([-1.2, -1.0, 0.1]), 0.16, 640, 480, (2.4, 2.4, 1.44), 8),
            ("plane_p1", E3, K1, np.array([0.0, -2.0, -1.0]), 0.5, 1220, 370, (4.0, 4.0, 2.0), 1)]


def gen_vox2pix():
    helpers = ref_import.data_helpers()
    out = {}
    for name, E, K, org, vs, W, H, scene, pid in vox2pix_cases():
        pix, fov, z = helpers.vox2pix(E, K, org, vs, W, H, scene, pid)
        out[name] = dict(cam_E=torch.from_numpy(E.copy()), cam_k=torch.from_numpy(K.copy()),
                         vox_origin=torch.from_numpy(org.copy()), voxel_size=vs, img_W=W, img_H=H,
                         scene_size=scene, pattern_id=pid, pix=torch.from_numpy(pix), fov=torch.from_numpy(fov),
                         pix_z=torch.from_numpy(z))
    torch.save(out, os.path.join(OUT, "vox2pix.pt"))


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = ref_import.modules()
    torch.manual_seed(0)
    gen_sfa(ref)
    gen_unet3d(ref)
    gen_occdepth(ref)
    gen_vox2pix()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
