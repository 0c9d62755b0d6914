"""Standalone config-2 lift (4 scales x 2 views, 128x128x16 voxels, 64 ch) for timing / ncu, in the precision mode
OCCDEPTH_PRECISION selects (tf32: fp32 features and output; bf16)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import synthetic as synth  # noqa: E402
from occdepth_b200.engine import CL  # noqa: E402
from occdepth_b200.models.SFA import lift_multiscale  # noqa: E402

H, W, C = 376, 1370, int(os.environ.get("LIFT_C", "64"))
FULL, PS = (256, 256, 32), 2
dev = torch.device("cuda")
PREC = os.environ.get("OCCDEPTH_PRECISION", "tf32")
TDT, ES = (torch.float32, 4) if PREC == "tf32" else (torch.bfloat16, 2)
pix, fov, _, _ = synth.kitti_indices(W, H, FULL, PS, voxel=0.2)
feats = []
U = 0
for s in (1, 2, 4, 8):
    h, w = synth.feature_hw(H, W, s)
    feats.append(torch.randn(2, h, w, C, device=dev).to(TDT))
    for v in range(2):
        idx = (pix[v, :, 0, 1] // s) * w + (pix[v, :, 0, 0] // s)
        U += int(torch.unique(idx[fov[v, :, 0]]).numel())
N = pix.shape[1]
out = CL.alloc(1, 128, 128, 16, C, dev, precision=PREC)
pix_d, fov_d = pix.to(dev), fov.to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
times = []
for i in range(13):
    flush.zero_()  # evict L2 between iterations (the feature maps alone are 131 MB)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    lift_multiscale(feats, [1, 2, 4, 8], pix_d, fov_d, out, "kitti", FULL, PS)
    e1.record()
    torch.cuda.synchronize()
    if i >= 3:
        times.append(e0.elapsed_time(e1))
ms = sorted(times)[len(times) // 2]
alg = U * C * ES + 2 * N * 17 + N * C * ES
print("lift (" + PREC + "): %.4f ms median; algorithmic bytes %.1f MB (U=%d) -> %.1f GB/s; fov %.3f" %
      (ms, alg / 1e6, U, alg / ms / 1e6, float(fov.float().mean())))
