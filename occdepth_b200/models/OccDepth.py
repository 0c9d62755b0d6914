"""Drop-in for the reference `occdepth/models/OccDepth.py` (class OccDepth, forward path :31-376).

Same constructor, same `forward(batch) -> dict`, same state_dict keys.  The forward is one planned sequence of
sm_100a launches: both stereo views go through the 2D UNet as one batch, the 4-scale / 2-view Stereo-SFA lift
is a single fused kernel writing the channels-last voxel grid the 3D UNet consumes, and the logits are written
in the reference's NCDHW fp32 layout by the last convolution's epilogue.
Training (`step`, losses, optimisers; OccDepth.py:378-600) is out of scope of this package.
"""
import os

import torch
import torch.nn as nn

from .. import _lib
from ..engine import CL, FnOp, Plan, require_cuda
from ._base import B200Module, _to_planar
from .SFA import SFA, lift_multiscale
from .unet2d import UNet2D
from .unet3d_kitti import UNet3D as UNet3DKitti
from .unet3d_nyu import UNet3D as UNet3DNYU

try:  # the reference derives from pl.LightningModule; use it when present so Trainer-based scripts keep working
    import pytorch_lightning as pl
    _Base = pl.LightningModule
except Exception:  # noqa: BLE001
    class _Base(nn.Module):
        def save_hyperparameters(self, *a, **k):
            pass

        def log(self, *a, **k):
            pass


def feature_hw(img_h, img_w, scale):
    """spatial size of the "1_scale" feature map (TF-SAME stride-2 stages: ceil division per halving)"""
    h, w, s = img_h, img_w, 1
    while s < scale:
        h, w, s = (h + 1) // 2, (w + 1) // 2, s * 2
    return h, w


def _get(config, name, default=None):
    try:
        return getattr(config, name)
    except (AttributeError, KeyError):
        return default


class OccDepth(_Base, B200Module):
    def __init__(self, class_names, class_weights, class_weights_occ=None, full_scene_size=None, project_res=[],
                 config=None, infer_mode=False):
        super().__init__()
        self.project_res = project_res
        self.full_scene_size = full_scene_size
        self.class_names = class_names
        self.class_weights = class_weights
        self.class_weights_occ = class_weights_occ
        self.dataset = config.dataset
        self.project_scale = config.project_scale
        self.n_relations = config.n_relations
        self.context_prior = config.context_prior
        self.n_classes = config.n_classes
        self.feature = config.feature
        self.feature_2d_oc = config.feature_2d_oc
        self.trans_2d_to_3d = config.trans_2d_to_3d
        self.cascade_cls = config.cascade_cls
        self.occluded_cls = config.occluded_cls
        self.multi_view_mode = _get(config, "multi_view_mode", True)
        self.share_2d_backbone_gradient = _get(config, "share_2d_backbone_gradient", False)
        print("INFO: Use cascade cls: {}".format(self.cascade_cls))
        print("INFO: Use occluded cls: {}".format(self.occluded_cls))
        self.infer_mode = infer_mode
        if self.infer_mode:
            self.context_prior = False
        self.use_stereo_depth_gt = _get(config, "use_stereo_depth_gt", False)
        self.use_lidar_depth_gt = _get(config, "use_lidar_depth_gt", False)
        self.use_depth_gt = _get(config, "use_depth_gt", False)
        assert not (self.use_stereo_depth_gt and self.use_lidar_depth_gt), "only with one depth data supported."
        self.with_depth_gt = self.use_stereo_depth_gt or self.use_lidar_depth_gt or self.use_depth_gt
        if self.dataset == "NYU":
            self.net_3d_decoder = UNet3DNYU(self.n_classes, nn.BatchNorm3d, n_relations=self.n_relations,
                                            feature=self.feature, full_scene_size=self.full_scene_size,
                                            context_prior=self.context_prior, cascade_cls=self.cascade_cls,
                                            infer_mode=self.infer_mode)
        elif self.dataset == "kitti":
            self.net_3d_decoder = UNet3DKitti(self.n_classes, nn.BatchNorm3d, project_scale=self.project_scale,
                                              feature=self.feature, full_scene_size=self.full_scene_size,
                                              context_prior=self.context_prior, cascade_cls=self.cascade_cls,
                                              occluded_cls=self.occluded_cls, infer_mode=self.infer_mode)
        self.net_rgb = UNet2D.build(out_feature=self.feature_2d_oc, use_decoder=True,
                                    backbone_2d_name=config.backbone_2d_name,
                                    return_up_feats=config.return_up_feats)
        self.save_hyperparameters()
        self.init_2d_to_3d_trans(config)
        if self.dataset not in ("kitti", "NYU"):
            raise NotImplementedError(self.dataset)

    def init_2d_to_3d_trans(self, config):
        print("INFO: Selected 2d->3d transformation method: {}".format(self.trans_2d_to_3d))
        if self.trans_2d_to_3d in ("flosp", "flosp_depth"):
            self.scale_2ds = [1, 2, 4, 8]
            self.projects = nn.ModuleDict({
                str(s): SFA(config.full_scene_size, project_scale=self.project_scale, dataset=self.dataset)
                for s in self.scale_2ds})
            if self.trans_2d_to_3d == "flosp_depth":
                from .flosp_depth.flosp_depth import FlospDepth, flosp_depth_conf_map
                self.flosp_depth_conf = flosp_depth_conf_map[self.dataset]
                self.flosp_depth_conf.update({
                    "scene_size": config.full_scene_size,
                    "project_scale": config.project_scale,
                    "output_channels": config.feature,
                    "depth_net_conf": dict(in_channels=config.feature,
                                           mid_channels=self.flosp_depth_conf["depth_net_conf"]["mid_channels"]),
                    "return_depth": self.with_depth_gt,
                    "infer_mode": self.infer_mode,
                })
                self.flosp_depth = FlospDepth(**self.flosp_depth_conf)
        else:
            raise NotImplementedError(f"{self.trans_2d_to_3d} is not supported yet.")

    # ------------------------------------------------------------------------------------------
    def _build(self, B, V, H, W, N, P, dev, batch):
        slab = self.__dict__.get("slab_ctx")
        plan = Plan(dev, slab=slab, precision=self.precision)
        ps = self.project_scale
        S = [int(s) // ps for s in self.full_scene_size]
        n_lo = 0
        if slab is not None:
            # one frame, X-slab partition: this rank lifts and decodes planes [x_lo, x_hi) of the voxel grid; the
            # 2D network is replicated (its receptive field is global: SE pooling) -- SURVEY.md section 8e
            if B != 1 or self.dataset != "kitti":
                raise NotImplementedError("slab partition: batch size 1, KITTI voxel order (X-major) only")
            x_lo, x_hi = slab.slab(S[0])
            n_lo, N = x_lo * S[1] * S[2], (x_hi - x_lo) * S[1] * S[2]
            S = [x_hi - x_lo, S[1], S[2]]
        virtual = V == 1 and "gt_depth" in batch                          # process_rgbs, OccDepth.py:221-229
        VL = 2 if virtual else V                                           # views seen by the lift
        # slab partition over an even number of ranks: the 2D network is sharded BY VIEW (its receptive field is
        # global inside a view -- SE pooling -- but the two views are independent, process_rgbs OccDepth.py:208-219):
        # ranks [0, R/2) run view 0, ranks [R/2, R) view 1, and one all-gather hands every rank both views' maps
        view_shard = (slab is not None and V == 2 and not virtual and slab.world % 2 == 0
                      and os.environ.get("OCCDEPTH_SLAB_VIEW_SHARD", "1") == "1")
        view_sel = (slab.rank // (slab.world // 2)) if view_shard else None
        img = plan.alloc(B * (1 if view_shard else V), 1, H, W, 3)
        Cf = self.feature_2d_oc
        assert Cf == self.feature, "feature_2d_oc must equal feature (the lift feeds the 3D net directly)"
        assert Cf % 8 == 0, "feature_2d_oc must be a multiple of 8"
        scales = [int(s) for s in self.project_res]
        views = {}
        outs = {}
        if virtual:
            # view-major buffers [2, B, h, w, C]: the 2D net writes view 0, the virtual-view kernel view 1
            for s in scales:
                h_s, w_s = feature_hw(H, W, s)
                vb = torch.zeros(2, B, h_s, w_s, Cf, dtype=plan.dtype, device=dev)
                views[s] = vb
                outs["1_%d" % s] = CL(vb[0].unsqueeze(1), Cf)
        gathered = None
        if view_shard:
            R2 = slab.world // 2
            hw = [feature_hw(H, W, s) for s in scales]
            offs, P_tot = [], 0
            for (h_s, w_s) in hw:
                offs.append(P_tot)
                P_tot += h_s * w_s
            slice_len = -(-P_tot // R2)
            P_pad = slice_len * R2
            packed = torch.zeros(P_pad, Cf, dtype=plan.dtype, device=dev)       # this rank's view, all scales
            gathered = torch.zeros(2, P_pad, Cf, dtype=plan.dtype, device=dev)  # both views after the all-gather
            for s, (h_s, w_s), o in zip(scales, hw, offs):
                outs["1_%d" % s] = CL(packed[o:o + h_s * w_s].view(1, 1, h_s, w_s, Cf), Cf)
        x_rgb = self.net_rgb.emit(plan, img, outs)                          # {"1_s": CL [B*V,1,h,w,Cf]}
        if view_shard:
            sl = slab.rank % R2
            plan.add(slab.all_gather_op(packed[sl * slice_len:(sl + 1) * slice_len],
                                        gathered.view(2 * R2, slice_len, Cf)))
            for s, (h_s, w_s), o in zip(scales, hw, offs):
                views[s] = gathered[:, o:o + h_s * w_s].view(2, 1, h_s, w_s, Cf)   # [view][B=1][h][w][C]
        depth0 = None
        if virtual:
            L = _lib.lib()
            gd = batch["gt_depth"]
            depth0 = torch.zeros(gd.shape[2], gd.shape[3], dtype=torch.float32, device=dev)
            bf = float(batch["virtual_bf"][0])
            for s in scales:
                vb = views[s]
                _, _, h_s, w_s, _ = vb.shape
                plan.add(FnOp(lambda st, vb=vb, h_s=h_s, w_s=w_s, s=s: L.occd_virtual_view_fwd(
                    vb[0].data_ptr(), vb[1].data_ptr(), depth0.data_ptr(), plan.lib_dtype, B, h_s, w_s, Cf, Cf, Cf,
                    depth0.shape[0], depth0.shape[1], bf / s, st), "virtual_view_1_%d" % s, keep=(vb, depth0)))
        pix = torch.zeros(B, VL, N, P, 2, dtype=torch.int64, device=dev)
        fov = torch.zeros(B, VL, N, P, dtype=torch.bool, device=dev)
        x3d = plan.alloc(B, S[0], S[1], S[2], self.feature)
        prior = None
        if self.trans_2d_to_3d == "flosp_depth":
            n_cams = 1 if self.dataset == "NYU" else V                      # OccDepth.py:303-305
            prior = self.flosp_depth.emit(plan, x_rgb, batch, B, V, n_cams)  # fp32 [B, X*Y*Z]
        for b in range(B):
            feats = []
            for s in scales:
                if virtual or view_shard:
                    feats.append(views[s][:, b])                            # [2, h, w, Cf], strided views
                else:
                    f = x_rgb["1_%d" % s]
                    assert f.coff == 0 and f.cstride == Cf
                    feats.append(f.buf[b * V:(b + 1) * V, 0])               # [V, h, w, Cf] contiguous view
            out_b = CL(x3d.buf[b:b + 1], x3d.C, 0, x3d.d0, x3d.dlen)
            pr = prior[b] if prior is not None else None
            plan.add(FnOp(lambda st, feats=feats, pb=pix[b], fb=fov[b], ob=out_b, pr=pr:
                          (lift_multiscale(feats, scales, pb, fb, ob, self.dataset,
                                           [S[0] * ps, S[1] * ps, S[2] * ps], ps,
                                           prior=pr, scale_const=100.0, stream=st), 0)[1],
                          "sfa_lift", keep=(feats, pix, fov, out_b)))
        out = self.net_3d_decoder.emit(plan, x3d)
        if slab is not None and self.trans_2d_to_3d == "flosp_depth":
            raise NotImplementedError("slab partition with the FlospDepth prior is not built")
        if os.environ.get("OCCDEPTH_CUDA_GRAPH", "1") == "1":
            try:
                plan.capture()
            except Exception as e:  # noqa: BLE001 -- same kernels either way; only the launch mechanism differs
                print("WARNING: CUDA graph capture failed (%r); launching kernels individually" % (e,))
                plan.graph = None
        return plan, img, pix, fov, out, depth0, n_lo, view_sel

    def forward(self, batch):
        img = batch["img"]
        if not img.is_cuda:
            if not torch.cuda.is_available():
                raise RuntimeError("OccDepth.forward: occdepth_b200 needs a CUDA (sm_100a) device")
            img = img.cuda(non_blocking=True)          # the reference moves the batch in forward too (:345)
        self._check_mode(img)
        dev = img.device
        if next(self.parameters()).device != dev:
            raise RuntimeError("OccDepth.forward: model and batch are on different devices")
        B, V, _, H, W = img.shape
        ps = self.project_scale
        pp = batch["projected_pix_{}".format(ps)]
        fm = batch["fov_mask_{}".format(ps)]
        N, P = pp[0].shape[1], pp[0].shape[2]
        virtual = V == 1 and "gt_depth" in batch
        slab = self.__dict__.get("slab_ctx")
        key = (B, V, H, W, N, P, str(dev), virtual, float(batch["virtual_bf"][0]) if virtual else 0.0,
               None if slab is None else (slab.rank, slab.world), self.precision, self.param_stamp())
        with torch.cuda.device(dev):           # launches, tensor maps and function attributes follow the tensors
            return self._forward_on(batch, key, img, pp, fm, B, V, H, W, N, P, dev)

    def prepare(self, batch):
        """build and cache the launch plan (weight snapshot, buffers, tensor maps, CUDA graph) for this batch's
        shapes without running it: the first forward() then costs what every later one does.  No collective is
        executed, so a multi-rank caller can agree on success before any rank starts a halo exchange."""
        self.__dict__["_build_only"] = True
        try:
            self.forward(batch)
        finally:
            self.__dict__.pop("_build_only", None)
        return self

    def _forward_on(self, batch, key, img, pp, fm, B, V, H, W, N, P, dev):
        ent = self._plans().get(key)
        if ent is None:
            self._plans().clear()              # weights changed / new shape: drop the stale snapshot
            with torch.no_grad():
                ent = self._build(B, V, H, W, N, P, dev, batch)
            self._plans()[key] = ent
        if self.__dict__.get("_build_only"):
            return None
        plan, img_cl, pix, fov, out, depth0, n_lo, view_sel = ent
        CL.from_planar(img.reshape(B * V, 3, H, W) if view_sel is None else img[:, view_sel].reshape(B, 3, H, W),
                       out=img_cl)
        if depth0 is not None:
            depth0.copy_(batch["gt_depth"][0, 0], non_blocking=True)
        nl = pix.shape[2]
        for b in range(B):
            pix[b].copy_(pp[b][:, n_lo:n_lo + nl], non_blocking=True)
            fov[b].copy_(fm[b][:, n_lo:n_lo + nl], non_blocking=True)
        if self.trans_2d_to_3d == "flosp_depth":
            self.flosp_depth.stage_inputs(batch, dev)
        plan.run()
        res = _to_planar(out, False)
        # tensors written directly by kernels live in plan-owned static buffers: hand out copies
        res = {k: (v.clone() if isinstance(out[k], torch.Tensor) else v) for k, v in res.items()}
        if self.with_depth_gt and self.trans_2d_to_3d == "flosp_depth":
            res["depth_pred"] = self.flosp_depth.depth_prob().clone()          # OccDepth.py:374-375
        return res

    @staticmethod
    def class_map(ssc_logit, inv_map=None):
        """uint16 class per voxel from the fp32 logits [B, C, X, Y, Z] on the GPU -- what the reference's callers
        compute on the host after copying all logits back (`np.argmax(torch.softmax(pred["ssc_logit"], 1).cpu(),
        1).astype(np.uint16)`, scripts/generate_output.py:94-97; eval.py does the same).  168 MB of logits shrink to
        a 4 MB map before the device->host read.  `inv_map` (int array of C label ids, io_data.get_inv_map) applies the
        submission writer's remap `inv_map[y_pred].astype(np.uint16)` (generate_kitti_submission.py:79) in the same
        launch."""
        from .. import _lib
        if not (ssc_logit.is_cuda and ssc_logit.dtype == torch.float32 and ssc_logit.dim() == 5):
            raise RuntimeError("OccDepth.class_map: expected CUDA fp32 logits [B, C, X, Y, Z]")
        x = ssc_logit.contiguous()
        B, Cn = x.shape[:2]
        S = x[0, 0].numel()
        out = torch.empty((B,) + tuple(x.shape[2:]), dtype=torch.uint16, device=x.device)
        lut = None
        if inv_map is not None:
            import numpy as np
            lut = np.ascontiguousarray(np.asarray(inv_map).astype(np.int32))
            if lut.shape != (Cn,):
                raise ValueError("OccDepth.class_map: inv_map must hold one label id per class")
        _lib.check(_lib.lib().occd_argmax_classes(x.data_ptr(), out.data_ptr(), B, Cn, S,
                                                  None if lut is None else lut.ctypes.data, _lib.stream_ptr()),
                   "occd_argmax_classes")
        return out

    def predict(self, batch, inv_map=None):
        """forward + class map: returns (y_pred uint16 [B, X, Y, Z], the forward's output dict)"""
        res = self.forward(batch)
        return self.class_map(res["ssc_logit"], inv_map), res

    def step(self, *a, **k):
        raise NotImplementedError("occdepth_b200 implements OccDepth.forward only (training is out of scope)")
