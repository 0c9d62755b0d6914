"""Stub of kornia 0.5.0 geometry helpers used by occdepth/models/f2v (see README.md)."""
import torch
import torch.nn.functional as F
from . import utils, geometry  # noqa: F401


def convert_points_to_homogeneous(points):
    return F.pad(points, [0, 1], "constant", 1.0)


def convert_points_from_homogeneous(points, eps=1e-8):
    z_vec = points[..., -1:]
    mask = torch.abs(z_vec) > eps
    scale = torch.where(mask, 1.0 / (z_vec + eps), torch.ones_like(z_vec))
    return scale * points[..., :-1]

from .geometry.linalg import transform_points  # noqa: E402,F401  (kornia 0.5.0 re-exports it at top level)
