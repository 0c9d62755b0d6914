"""3D network drop-in modules on CUDA vs the CPU fp32 oracle (same seeded weights)."""
import pytest
import torch
import torch.nn as nn

from oracle import functional as OF
from oracle import synth

pytestmark = pytest.mark.gpu

# bf16 operands / bf16 activations between ~25 fused layers, fp32 accumulation: stated tolerance on
# max-abs-diff relative to max-abs of the oracle tensor
TOL = 4e-2


def _check(got, want, tol=TOL):
    for k in want:
        g, w = got[k].float().cpu(), want[k]
        assert g.shape == w.shape, (k, g.shape, w.shape)
        err = float((g - w).abs().max() / w.abs().max().clamp_min(1e-6))
        assert err <= tol, (k, err)


def test_process_downsample_upsample():
    from occdepth_b200.models.modules import Downsample, Process, Upsample
    torch.manual_seed(0)
    proc = Process(32, nn.BatchNorm3d, 0.1).eval()
    down = Downsample(32, nn.BatchNorm3d, 0.1).eval()
    up = Upsample(64, 32, nn.BatchNorm3d, 0.1).eval()
    for m in (proc, down, up):
        synth.randomize_bn_(m)
    x = torch.randn(1, 32, 12, 10, 8)
    with torch.no_grad():
        want_p = OF.process({"p." + k: v for k, v in proc.state_dict().items()}, "p", x)
        want_d = OF.downsample({"p." + k: v for k, v in down.state_dict().items()}, "p", x)
        want_u = OF.upsample({"p." + k: v for k, v in up.state_dict().items()}, "p", want_d)
        _check({"p": proc.cuda()(x.cuda())}, {"p": want_p}, 2e-2)
        _check({"d": down.cuda()(x.cuda())}, {"d": want_d}, 2e-2)
        _check({"u": up.cuda()(want_d.cuda())}, {"u": want_u}, 2e-2)


@pytest.mark.parametrize("which", ["kitti", "nyu"])
def test_unet3d(which):
    torch.manual_seed(0)
    if which == "kitti":
        from occdepth_b200.models.unet3d_kitti import UNet3D
        full, ps, f = (32, 32, 16), 2, 32
        m = UNet3D(20, nn.BatchNorm3d, full, f, ps, context_prior=True, cascade_cls=True, occluded_cls=True).eval()
        x = torch.randn(1, f, 16, 16, 8)
    else:
        from occdepth_b200.models.unet3d_nyu import UNet3D
        full, f = (12, 8, 12), 24
        m = UNet3D(12, nn.BatchNorm3d, f, full, context_prior=True, cascade_cls=False).eval()
        x = torch.randn(1, f, 12, 8, 12)
    synth.randomize_bn_(m)
    sd = {"n." + k: v.clone() for k, v in m.state_dict().items()}
    with torch.no_grad():
        if which == "kitti":
            want = OF.unet3d_kitti(sd, "n", x, full, ps, True, True, True)
        else:
            want = OF.unet3d_nyu(sd, "n", x, full, 4, True, False)
        got = m.cuda()({"x3d": x.cuda()})
    assert set(got.keys()) == set(want.keys())
    _check(got, want)
    agree = (got["ssc_logit"].argmax(1).cpu() == want["ssc_logit"].argmax(1)).float().mean()
    assert agree > 0.9, float(agree)
