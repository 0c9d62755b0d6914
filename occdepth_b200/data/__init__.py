"""Device-side pieces of the reference's data pipeline (`occdepth.data`), SURVEY 8f row 2.

`vox2pix` mirrors `occdepth.data.utils.helpers.vox2pix` (helpers.py:94-169): same arguments, same three results
(`projected_pix` int64 (N, P, 2), `fov_mask` bool (N, P), `pix_z` (N,) in the pose's precision), bit for bit -- but computed by
`occd_vox2pix_fwd` on the GPU and returned as CUDA tensors, so the 8.4 MB of indices per stereo frame never
cross PCIe and the per-sample numba job disappears from the loader.
"""
import ctypes as C

import numpy as np
import torch

from .. import _lib

# fusion.py:238-330 `pixel_partern` ("dso residual pattern"): (dx, dy) offsets around the projected centre
PIXEL_PATTERNS = [
    [[0, 0]],
    [[0, 0], [0, -1], [-1, 0], [1, 0], [0, 1]],
    [[0, 0], [-1, -1], [1, 1], [-1, 1], [1, -1]],
    [[0, 0], [-1, -1], [-1, 0], [-1, 1], [-1, 0], [0, 1], [1, -1], [1, 0], [1, 1]],
    [[0, 0], [0, -2], [-1, -1], [1, -1], [-2, 0], [2, 0], [-1, 1], [1, 1], [0, 2]],
    [[0, 0], [0, -2], [-1, -1], [1, -1], [-2, 0], [2, 0], [-1, 1], [1, 1], [0, 2], [-2, -2], [-2, 2], [2, -2],
     [2, 2]],
    [[0, 0]] + [[a, b] for a in (-2, -1, 0, 1, 2) for b in (-2, -1, 0, 1, 2) if (a, b) != (0, 0)],
    [[0, 0], [0, -2], [-1, -1], [1, -1], [-2, 0], [2, 0], [-1, 1], [0, 2]],
    [[0, 0], [0, -2], [-1, -1], [1, -1], [-2, 0], [2, 0], [-1, 1], [1, 1], [0, 2], [-2, -2], [-2, 2], [2, -2],
     [2, 2], [-3, -1], [-3, 1], [3, -1], [3, 1], [1, -3], [-1, -3], [1, 3], [-1, 3]],
]


def volume_dims(vox_origin, voxel_size, scene_size):
    """helpers.py:126-134, the same numpy expression (so the same rounding of the extent)"""
    vol_bnds = np.zeros((3, 2))
    vol_bnds[:, 0] = vox_origin
    vol_bnds[:, 1] = np.asarray(vox_origin) + np.array(scene_size)
    return np.ceil((vol_bnds[:, 1] - vol_bnds[:, 0]) / voxel_size).copy(order="C").astype(int)


def vox2pix(cam_E, cam_k, vox_origin, voxel_size, img_W, img_H, scene_size, pattern_id, device=None):
    """occdepth.data.utils.helpers.vox2pix on the GPU (see module docstring).  Raises if no CUDA device."""
    dev = torch.device("cuda" if device is None else device)
    if dev.type != "cuda" or not torch.cuda.is_available():
        raise RuntimeError("occdepth_b200.data.vox2pix runs on CUDA (sm_100a) only -- there is no CPU fallback")
    if not 0 <= int(pattern_id) < len(PIXEL_PATTERNS):
        raise IndexError("pattern_id out of range")        # the reference indexes the same 9-entry list
    X, Y, Z = (int(v) for v in volume_dims(vox_origin, voxel_size, scene_size))
    E = np.asarray(cam_E)
    f32_pose = E.dtype == np.float32          # np.dot(cam_E, [pts 1]) stays in float32 for a float32 pose
    E = np.ascontiguousarray(E.astype(np.float32 if f32_pose else np.float64))
    k32 = np.ascontiguousarray(np.asarray(cam_k).astype(np.float32))         # fusion.py:331 intr.astype(float32)
    o32 = np.ascontiguousarray(np.asarray(vox_origin).astype(np.float32))    # fusion.py:205
    if E.shape != (4, 4) or k32.shape != (3, 3) or o32.shape != (3,):
        raise ValueError("vox2pix: cam_E must be 4x4, cam_k 3x3, vox_origin (3,)")
    pat = np.ascontiguousarray(np.asarray(PIXEL_PATTERNS[int(pattern_id)], dtype=np.int32))
    P, N = len(pat), X * Y * Z
    with torch.cuda.device(dev):
        pix = torch.empty(N, P, 2, dtype=torch.int64, device=dev)
        fov = torch.empty(N, P, dtype=torch.bool, device=dev)
        pix_z = torch.empty(N, dtype=torch.float32 if f32_pose else torch.float64, device=dev)
        rc = _lib.lib().occd_vox2pix_fwd(E.ctypes.data_as(C.c_void_p), 1 if f32_pose else 0,
                                         k32.ctypes.data_as(C.c_void_p),
                                         o32.ctypes.data_as(C.c_void_p), float(voxel_size), X, Y, Z, int(img_W),
                                         int(img_H), pat.ctypes.data_as(C.c_void_p), P, pix.data_ptr(),
                                         fov.data_ptr(), pix_z.data_ptr(), _lib.stream_ptr())
        _lib.check(rc, "occd_vox2pix_fwd")
    return pix, fov, pix_z


IMAGENET_MEAN, IMAGENET_STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)     # kitti_dataset.py:167-169


def normalize_rgb(img_u8, img_H, img_W, mean=IMAGENET_MEAN, std=IMAGENET_STD, device=None):
    """The datasets' image pre-processing on the GPU: uint8 RGB (H0, W0, 3) array / tensor (host or device) ->
    float32 (3, img_H, img_W) CUDA tensor equal, bit for bit, to
    `normalize_rgb((np.array(img, np.float32) / 255.0)[:img_H, :img_W])` of kitti_dataset.py:376-402
    (ToTensor + Normalize).  The host->device copy is the uint8 image (4x smaller than the float tensor)."""
    dev = torch.device("cuda" if device is None else device)
    if dev.type != "cuda" or not torch.cuda.is_available():
        raise RuntimeError("occdepth_b200.data.normalize_rgb runs on CUDA (sm_100a) only -- there is no CPU fallback")
    t = torch.as_tensor(np.asarray(img_u8) if not isinstance(img_u8, torch.Tensor) else img_u8)
    if t.dtype != torch.uint8 or t.dim() != 3 or t.shape[2] != 3:
        raise ValueError("normalize_rgb: expected a uint8 (H, W, 3) RGB image")
    t = t.contiguous().to(dev, non_blocking=True)
    H0, W0 = int(t.shape[0]), int(t.shape[1])
    m32 = np.ascontiguousarray(np.asarray(mean, dtype=np.float32))
    s32 = np.ascontiguousarray(np.asarray(std, dtype=np.float32))
    with torch.cuda.device(dev):
        out = torch.empty(3, int(img_H), int(img_W), dtype=torch.float32, device=dev)
        rc = _lib.lib().occd_normalize_rgb_u8(t.data_ptr(), out.data_ptr(), H0, W0, int(img_H), int(img_W),
                                              m32.ctypes.data_as(C.c_void_p), s32.ctypes.data_as(C.c_void_p),
                                              _lib.stream_ptr())
        _lib.check(rc, "occd_normalize_rgb_u8")
    return out
