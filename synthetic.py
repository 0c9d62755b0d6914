"""Seeded synthetic inputs / weights for the bench harness and the parity tests (harness code: neither part of the
product package nor of the oracle -- it only manufactures inputs).

Follows SURVEY.md section 8d: `torch.randn` images, module-default init under a fixed seed followed by
randomised BatchNorm statistics/affine (so that BN folding is exercised and logits are O(1)), and index
tensors from a restatement of the reference's `vox2pix` (occdepth/data/utils/helpers.py:94-169,
fusion.py:201-232) with a KITTI-like calibration scaled to the synthetic image size.
"""
import numpy as np
import torch


def randomize_bn_(sd_or_module, seed=1):
    """running_mean ~ N(0,0.1), running_var ~ U(0.5,1.5), weight ~ U(0.5,1.5), bias ~ N(0,0.1)."""
    g = torch.Generator().manual_seed(seed)
    sd = sd_or_module.state_dict() if hasattr(sd_or_module, "state_dict") else sd_or_module
    for k in sorted(sd.keys()):
        if k.endswith("running_mean"):
            base = k[: -len("running_mean")]
            n = sd[k].numel()
            sd[base + "running_mean"].copy_(torch.randn(n, generator=g) * 0.1)
            sd[base + "running_var"].copy_(torch.rand(n, generator=g) + 0.5)
            sd[base + "weight"].copy_(torch.rand(n, generator=g) + 0.5)
            sd[base + "bias"].copy_(torch.randn(n, generator=g) * 0.1)
    return sd


def kitti_calib(img_w, img_h, focal=None):
    """KITTI-like intrinsics / lidar->camera extrinsics for the two stereo views.
    focal defaults to 707.0912 px at the 1370-px-wide synthetic image and scales with the image width."""
    f = 707.0912 * img_w / 1370.0 if focal is None else focal
    K = np.array([[f, 0, img_w / 2.0], [0, f, img_h / 2.0], [0, 0, 1]], dtype=np.float64)
    T0 = np.array([[0, -1, 0, 0], [0, 0, -1, -0.08], [1, 0, 0, -0.27], [0, 0, 0, 1]], dtype=np.float32)
    T1 = T0.copy()
    T1[0, 3] -= 0.54
    return K, [T0, T1]


def vox2pix(cam_E, cam_k, vox_origin, voxel_size, img_W, img_H, scene_size):
    """helpers.py:94-169 for pattern_id 0 (P = 1).  Returns projected_pix (N,1,2) int64 (x,y), fov (N,1) bool,
    pix_z (N,).  Voxel order: C-order over (X,Y,Z) (helpers.py:137-146)."""
    vox_origin = np.asarray(vox_origin, dtype=np.float64)
    vol_dim = np.ceil(np.asarray(scene_size, dtype=np.float64) / voxel_size).astype(int)
    xv, yv, zv = np.meshgrid(range(vol_dim[0]), range(vol_dim[1]), range(vol_dim[2]), indexing="ij")
    vox = np.stack([xv.reshape(-1), yv.reshape(-1), zv.reshape(-1)], 1).astype(np.float32)
    # fusion.py:201-217 vox2world: float32 origin / coordinates, float64 scalar voxel size, stored as float32
    vo = vox_origin.astype(np.float32)
    pts = (vo[None, :].astype(np.float64) + voxel_size * vox.astype(np.float64) + voxel_size * 0.5).astype(np.float32)
    # fusion.py rigid_transform: [pts 1] @ E^T
    E = np.asarray(cam_E)
    pts_h = np.hstack([pts, np.ones((len(pts), 1), dtype=np.float32)])
    cam = (E @ pts_h.T).T[:, :3]
    k = np.asarray(cam_k).astype(np.float32)
    cam32 = cam.astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        px = np.round(cam32[:, 0] * k[0, 0] / cam32[:, 2] + k[0, 2])
        py = np.round(cam32[:, 1] * k[1, 1] / cam32[:, 2] + k[1, 2])
    px = np.nan_to_num(px, nan=-1e6, posinf=1e9, neginf=-1e9).astype(np.int64)
    py = np.nan_to_num(py, nan=-1e6, posinf=1e9, neginf=-1e9).astype(np.int64)
    pz = cam32[:, 2]
    fov = (px >= 0) & (px < img_W) & (py >= 0) & (py < img_H) & (pz > 0)
    pix = np.stack([px, py], 1)[:, None, :]
    return torch.from_numpy(pix), torch.from_numpy(fov[:, None]), pz


def kitti_indices(img_w, img_h, full_scene_size=(256, 256, 32), project_scale=2, voxel=0.2, n_views=2):
    """projected_pix (V,N,1,2) int64 and fov_mask (V,N,1) bool for the lift grid of `project_scale`."""
    K, Ts = kitti_calib(img_w, img_h)
    scene_m = tuple(s * voxel for s in full_scene_size)
    pix, fov = [], []
    for v in range(n_views):
        p, f, _ = vox2pix(Ts[v], K, (0.0, -scene_m[1] / 2.0, -2.0), voxel * project_scale, img_w, img_h, scene_m)
        pix.append(p)
        fov.append(f)
    return torch.stack(pix), torch.stack(fov), K, Ts


def random_indices(n, img_w, img_h, n_views=2, P=1, seed=0, margin=(40, 20)):
    """uniform-random pixels incl. out-of-image ones (plumbing tests only; destroys locality)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randint(-margin[0], img_w + margin[0], (n_views, n, P), generator=g)
    y = torch.randint(-margin[1], img_h + margin[1], (n_views, n, P), generator=g)
    fov = (x >= 0) & (x < img_w) & (y >= 0) & (y < img_h)
    return torch.stack([x, y], -1).long(), fov


def feature_hw(img_h, img_w, scale):
    """spatial size of the "1_scale" feature map: TF-SAME stride-2 stages -> ceil division."""
    h, w = img_h, img_w
    s = 1
    while s < scale:
        h, w = (h + 1) // 2, (w + 1) // 2
        s *= 2
    return h, w


class Cfg(dict):
    """attribute-style config object accepted by the reference OccDepth.__init__ (OccDepth.py:49-96)."""
    __getattr__ = dict.__getitem__


def occdepth_cfg(**over):
    c = Cfg(dataset="kitti", frustum_size=8, project_scale=2, n_relations=4, lr=2e-4, weight_decay=1e-4,
            fp_loss=True, context_prior=True, relation_loss=True, CE_ssc_loss=True, sem_scal_loss=True,
            geo_scal_loss=True, n_classes=20, feature=64, feature_2d_oc=64, trans_2d_to_3d="flosp",
            cascade_cls=True, occluded_cls=False, sem_step_decay_loss=False, multi_view_mode=True,
            share_2d_backbone_gradient=False, use_stereo_depth_gt=False, use_lidar_depth_gt=False,
            use_depth_gt=False, depth_loss_weight=1.0, backbone_2d_name="tf_efficientnet_b7_ns",
            return_up_feats=1, batch_size_per_gpu=1, n_gpus=1, full_scene_size=(256, 256, 32))
    c.update(over)
    return c


def seed_weights_(module, seed=0):
    """Deterministic weights independent of module construction order / torch init code: every state_dict entry is
    filled, in sorted key order, from one seeded CPU generator (conv/linear weights ~ N(0, 1/fan_in), biases ~
    N(0, 0.1), BatchNorm weight/var ~ U(0.5, 1.5), mean ~ N(0, 0.1)).  The same call reproduces the same weights
    for the reference module, the oracle and the CUDA modules (identical key sets)."""
    g = torch.Generator().manual_seed(seed)
    sd = module.state_dict()
    for k in sorted(sd.keys()):
        t = sd[k]
        if not t.dtype.is_floating_point:
            continue
        n = t.numel()
        if k.endswith("running_var"):
            v = torch.rand(n, generator=g) + 0.5
        elif k.endswith("running_mean"):
            v = torch.randn(n, generator=g) * 0.1
        elif t.dim() <= 1:
            is_bn_weight = k.endswith("weight") and (k[: -len("weight")] + "running_var") in sd
            v = torch.rand(n, generator=g) + 0.5 if is_bn_weight else torch.randn(n, generator=g) * 0.1
        else:
            fan_in = max(1, n // t.shape[0])
            if "ConvTranspose" in k:
                fan_in = max(1, n // t.shape[1])
            v = torch.randn(n, generator=g) / fan_in ** 0.5
        t.copy_(v.view(t.shape).to(t.dtype))
    return module
