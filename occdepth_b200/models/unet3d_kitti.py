"""Drop-in for the reference `occdepth/models/unet3d_kitti.py` (UNet3D :14-126)."""
import torch
import torch.nn as nn

from ._base import B200Module
from .CRP3D import CPMegaVoxels
from .modules import (Convblock3d, Downsample, Process, SegmentationHead, SegmentationHeadCascadeCLS,
                      SegmentationHeadOccludedCLS, Upsample)


class UNet3D(B200Module):
    def __init__(self, class_num, norm_layer, full_scene_size, feature, project_scale, context_prior=None,
                 bn_momentum=0.1, cascade_cls=False, occluded_cls=False, infer_mode=False):
        super(UNet3D, self).__init__()
        self.business_layer = []
        self.project_scale = project_scale
        self.full_scene_size = full_scene_size
        self.feature = feature
        self.cascade_cls = cascade_cls
        self.occluded_cls = occluded_cls
        self.infer_mode = infer_mode
        size_l1 = (int(self.full_scene_size[0] / project_scale), int(self.full_scene_size[1] / project_scale),
                   int(self.full_scene_size[2] / project_scale))
        size_l2 = (size_l1[0] // 2, size_l1[1] // 2, size_l1[2] // 2)
        size_l3 = (size_l2[0] // 2, size_l2[1] // 2, size_l2[2] // 2)
        dilations = [1, 2, 3]
        self.process_l1 = nn.Sequential(
            Process(self.feature, norm_layer, bn_momentum, dilations=[1, 2, 3]),
            Downsample(self.feature, norm_layer, bn_momentum))
        self.process_l2 = nn.Sequential(
            Process(self.feature * 2, norm_layer, bn_momentum, dilations=[1, 2, 3]),
            Downsample(self.feature * 2, norm_layer, bn_momentum))
        self.up_13_l2 = Upsample(self.feature * 4, self.feature * 2, norm_layer, bn_momentum)
        self.up_12_l1 = Upsample(self.feature * 2, self.feature, norm_layer, bn_momentum)
        if self.project_scale == 1:
            self.up_l1_lfull = Convblock3d(self.feature, self.feature // 2, norm_layer, bn_momentum, stride=1)
        else:
            self.up_l1_lfull = Upsample(self.feature, self.feature // 2, norm_layer, bn_momentum)
        if self.cascade_cls:
            self.ssc_head = SegmentationHeadCascadeCLS(self.feature // 2, self.feature // 2, class_num, dilations)
        else:
            self.ssc_head = SegmentationHead(self.feature // 2, self.feature // 2, class_num, dilations)
        if self.occluded_cls:
            self.occluded_head = SegmentationHeadOccludedCLS(self.feature // 2, self.feature // 2, class_num,
                                                             dilations)
        self.context_prior = context_prior
        if context_prior:
            self.CP_mega_voxels = CPMegaVoxels(self.feature * 4, size_l3, bn_momentum=bn_momentum)

    def emit(self, plan, x3d_l1):
        res = {}
        x3d_l2 = self.process_l1[1].emit(plan, self.process_l1[0].emit(plan, x3d_l1))
        x3d_l3 = self.process_l2[1].emit(plan, self.process_l2[0].emit(plan, x3d_l2))
        if self.context_prior:
            ret = self.CP_mega_voxels.emit(plan, x3d_l3)
            x3d_l3 = ret["x"]
            for k in ret.keys():
                res[k] = ret[k]
        x3d_up_l2 = self.up_13_l2.emit(plan, x3d_l3, skip=x3d_l2)
        x3d_up_l1 = self.up_12_l1.emit(plan, x3d_up_l2, skip=x3d_l1)
        x3d_up_lfull = self.up_l1_lfull.emit(plan, x3d_up_l1)
        if not self.infer_mode:
            res["x3d_l1"] = x3d_up_l1
            res["x3d_l2"] = x3d_up_l2
            res["x3d_l3"] = x3d_l3
        if self.cascade_cls:
            ssc_logit_full, ssc_logit_full_occ = self.ssc_head.emit(plan, x3d_up_lfull)
            res["ssc_logit"] = ssc_logit_full
            if not self.infer_mode:
                res["occ_logit"] = ssc_logit_full_occ
        else:
            res["ssc_logit"] = self.ssc_head.emit(plan, x3d_up_lfull)
        if self.occluded_cls:
            occluded_logit_full = self.occluded_head.emit(plan, x3d_up_lfull)
            if not self.infer_mode:
                res["occluded_logit"] = occluded_logit_full
        return res

    def forward(self, input_dict):
        return self._run_planar(input_dict["x3d"])
