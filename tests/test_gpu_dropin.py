"""The reference's caller loop against the drop-in (occdepth/scripts/generate_output.py:73-97): the model class is
imported through the reference's own import path after `install_as_occdepth()`, a reference-layout state_dict is
loaded with strict=True, and the script's loop body runs unchanged."""
import numpy as np
import pytest
import torch

from oracle import functional as OF
from oracle import ref_import, synth

pytestmark = pytest.mark.gpu


def to_cuda(datas):                      # generate_output.py:16-19, verbatim semantics
    assert isinstance(datas, list)
    for i, data in enumerate(datas):
        datas[i] = data.cuda()


def test_generate_output_loop_body():
    import occdepth_b200
    occdepth_b200.install_as_occdepth()
    from occdepth.models.OccDepth import OccDepth          # generate_output.py:2
    assert OccDepth.__module__ == "occdepth_b200.models.OccDepth"
    torch.manual_seed(0)
    full, ps, ncls = (32, 32, 16), 2, 20
    cfg = synth.occdepth_cfg(dataset="kitti", full_scene_size=full, project_scale=ps, feature=32, feature_2d_oc=32,
                             n_classes=ncls, cascade_cls=True, backbone_2d_name="tf_efficientnet_b3_ns")
    with ref_import.quiet():
        donor = OccDepth(["c"] * ncls, torch.ones(ncls), full_scene_size=full, project_res=["1", "2", "4", "8"],
                         config=cfg)
        synth.randomize_bn_(donor)
        sd = {k: v.clone() for k, v in donor.state_dict().items()}
        model = OccDepth(["c"] * ncls, torch.ones(ncls), full_scene_size=full, project_res=["1", "2", "4", "8"],
                         config=cfg)
    model.load_state_dict(sd, strict=True)                 # load_from_checkpoint(..., strict=True), :73-78
    model.cuda()
    model.eval()                                           # :79-80
    H, W = 33, 49
    g = torch.Generator().manual_seed(0)
    N = (full[0] // ps) * (full[1] // ps) * (full[2] // ps)
    pix, fov = synth.random_indices(N, W, H, n_views=2, P=1, seed=5, margin=(10, 6))
    batch = {"img": torch.randn(1, 2, 3, H, W, generator=g), "projected_pix_2": [pix], "fov_mask_2": [fov],
             "T_velo_2_cam": [torch.eye(4).repeat(2, 1, 1)], "cam_k": [torch.eye(3, dtype=torch.float64).repeat(2, 1, 1)],
             "ida_mats": [torch.eye(4).repeat(2, 1, 1)]}
    ocfg = dict(cfg)
    ocfg["project_res"] = ["1", "2", "4", "8"]
    with torch.no_grad():
        want = OF.occdepth_forward(sd, {"img": batch["img"], "projected_pix_2": [pix], "fov_mask_2": [fov]}, ocfg)
        # ---- loop body, generate_output.py:88-97 ----
        batch["img"] = batch["img"].cuda()
        to_cuda(batch["T_velo_2_cam"])
        to_cuda(batch["cam_k"])
        to_cuda(batch["ida_mats"])
        pred = model(batch)
        y_pred = torch.softmax(pred["ssc_logit"], dim=1).detach().cpu().numpy()
        y_pred = np.argmax(y_pred, axis=1)
        out_dict = {"y_pred": y_pred[0].astype(np.uint16)}
    y_want = want["ssc_logit"].argmax(1)[0].numpy().astype(np.uint16)
    assert out_dict["y_pred"].shape == y_want.shape == full
    assert float((out_dict["y_pred"] == y_want).mean()) >= 0.995
    # the device-side post-processing returns the same map the script computes on the host
    assert np.array_equal(model.class_map(pred["ssc_logit"]).cpu().numpy()[0], out_dict["y_pred"])
    # a weight edited in place (optimizer step, param.copy_(), load_state_dict on a CHILD module) is picked up by the
    # next forward: plans are keyed on the parameters' storage pointers and in-place version counters
    with torch.no_grad():
        model.net_3d_decoder.ssc_head.conv_classes.bias.add_(1000.0 * torch.eye(ncls, device="cuda")[3])
        pred2 = model(batch)
    assert float((pred2["ssc_logit"].argmax(1) == 3).float().mean()) == 1.0
    child_sd = {k: v.clone() for k, v in model.net_3d_decoder.state_dict().items()}
    child_sd["ssc_head.conv_classes.bias"] = child_sd["ssc_head.conv_classes.bias"] - 1000.0 * torch.eye(ncls, device="cuda")[3] \
        + 1000.0 * torch.eye(ncls, device="cuda")[5]
    model.net_3d_decoder.load_state_dict(child_sd, strict=True)
    with torch.no_grad():
        pred3 = model(batch)
    assert float((pred3["ssc_logit"].argmax(1) == 5).float().mean()) == 1.0


def test_frame_pipeline_matches_forward():
    """occdepth_b200.serving.FramePipeline (copy streams + double buffers around forward): every frame's read-back
    equals what a plain forward returns for that frame"""
    from occdepth_b200.models.OccDepth import OccDepth
    from occdepth_b200.serving import FramePipeline
    torch.manual_seed(0)
    full, ps, ncls = (32, 32, 16), 2, 20
    cfg = synth.occdepth_cfg(dataset="kitti", full_scene_size=full, project_scale=ps, feature=32, feature_2d_oc=32,
                             n_classes=ncls, cascade_cls=True, backbone_2d_name="tf_efficientnet_b3_ns")
    with ref_import.quiet():
        model = OccDepth(["c"] * ncls, torch.ones(ncls), full_scene_size=full, project_res=["1", "2", "4", "8"],
                         config=cfg)
        synth.randomize_bn_(model)
    model.cuda().eval()
    H, W = 33, 49
    N = (full[0] // ps) * (full[1] // ps) * (full[2] // ps)
    frames = []
    for i in range(5):
        g = torch.Generator().manual_seed(i)
        pix, fov = synth.random_indices(N, W, H, n_views=2, P=1, seed=5 + i, margin=(10, 6))
        frames.append((torch.randn(1, 2, 3, H, W, generator=g).pin_memory(), pix.pin_memory(), fov.pin_memory()))
    with torch.no_grad():
        want = [model({"img": f[0].cuda(), "projected_pix_2": [f[1]], "fov_mask_2": [f[2]]})["ssc_logit"].cpu()
                for f in frames]
    pipe = FramePipeline(model, frames[0][0].shape, frames[0][1].shape, frames[0][2].shape)
    got = {}
    for f in frames:
        t = pipe.submit(*f)
        if t is not None:
            got[t] = pipe.result(t).clone()
    got[pipe.flush()] = pipe.result(pipe.flush()).clone()
    assert sorted(got) == list(range(5))
    for i in range(5):
        assert torch.equal(got[i], want[i]), i
    with pytest.raises(ValueError):
        pipe.result(0)
