"""X-slab partition of one frame (BASELINE.json configs[2]): R simulated ranks on one GPU, executed in lock-step
(parallel.SimSlabGroup), must reproduce the un-partitioned CUDA result; the NCCL flavour of the same exchange ops is
covered by tests/test_parallel_gloo.py (gloo, CPU) and tools/slab_check.py (torchrun, N GPUs)."""
import pytest
import torch
import torch.nn as nn

from oracle import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2, 4])
def test_unet3d_slab_matches_unpartitioned(world):
    from occdepth_b200.engine import CL
    from occdepth_b200.models.unet3d_kitti import UNet3D
    from occdepth_b200.parallel import SimSlabGroup
    torch.manual_seed(0)
    full, ps, f = (128, 32, 16), 2, 32       # l1 64x16x8, l3 16x4x2: 4 ranks -> 4 planes per rank at l3 (>= dil 3)
    m = UNet3D(20, nn.BatchNorm3d, full, f, ps, context_prior=True, cascade_cls=True).eval()
    synth.seed_weights_(m, 4)
    m = m.cuda()
    x = torch.randn(1, f, 64, 16, 8).cuda()
    with torch.no_grad():
        ref = m({"x3d": x})
        grp = SimSlabGroup(world, halo=3)
        ents = []
        per = 64 // world
        for r, ctx in enumerate(grp.ctxs):
            m.enable_slab_parallel(ctx)
            xl = x[:, :, r * per:(r + 1) * per].contiguous()
            plan, xin, y = m._get_plan(xl)
            CL.from_planar(xl, out=xin)
            ents.append((plan, y))
        SimSlabGroup.run_lockstep([e[0] for e in ents])
        torch.cuda.synchronize()
        m.__dict__.pop("slab_ctx")
    assert grp.ctxs[0].n_exchanges >= 20     # halo exchanges were actually planned
    from occdepth_b200.models._base import _to_planar
    outs = [_to_planar(y, False) for _, y in ents]
    for k in ("ssc_logit", "occ_logit", "x3d_l1", "x3d_l2", "x3d_l3"):
        got = torch.cat([o[k] for o in outs], 2)
        assert got.shape == ref[k].shape, k
        err = float((got - ref[k]).abs().max() / ref[k].abs().max())
        assert err <= 1e-3, (k, err)
    got = torch.cat([o["P_logits"] for o in outs], 3)
    assert float((got - ref["P_logits"]).abs().max() / ref["P_logits"].abs().max()) <= 1e-3
