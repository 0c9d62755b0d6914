"""Drop-in for the reference `occdepth/models/DDR.py`: Bottleneck3D (:35-139).

Same constructor / parameters / state_dict keys; forward = 5 (stride 1) or 9 (stride 2) fused implicit-GEMM
launches: BatchNorm folded, ReLU and the cross-stage adds in the epilogues, AvgPool3d+1x1x1 side paths as
strided multi-tap convolutions.
"""
import torch
import torch.nn as nn

from ..engine import fold_bn
from ._base import B200Module


def _pool_conv_weight(w, k):
    """AvgPool3d(kernel=k, stride=k) followed by a 1x1x1 conv == conv with kernel k, stride k, weights w/|k|."""
    n = k[0] * k[1] * k[2]
    return (w / n).expand(-1, -1, k[0], k[1], k[2]).contiguous()


class Bottleneck3D(B200Module):
    def __init__(self, inplanes, planes, norm_layer, stride=1, dilation=[1, 1, 1], expansion=4, downsample=None,
                 fist_dilation=1, multi_grid=1, bn_momentum=0.0003):
        super(Bottleneck3D, self).__init__()
        self.expansion = expansion
        self.conv1 = nn.Conv3d(inplanes, planes, kernel_size=1, bias=False)
        self.bn1 = norm_layer(planes, momentum=bn_momentum)
        self.conv2 = nn.Conv3d(planes, planes, kernel_size=(1, 1, 3), stride=(1, 1, stride),
                               dilation=(1, 1, dilation[0]), padding=(0, 0, dilation[0]), bias=False)
        self.bn2 = norm_layer(planes, momentum=bn_momentum)
        self.conv3 = nn.Conv3d(planes, planes, kernel_size=(1, 3, 1), stride=(1, stride, 1),
                               dilation=(1, dilation[1], 1), padding=(0, dilation[1], 0), bias=False)
        self.bn3 = norm_layer(planes, momentum=bn_momentum)
        self.conv4 = nn.Conv3d(planes, planes, kernel_size=(3, 1, 1), stride=(stride, 1, 1),
                               dilation=(dilation[2], 1, 1), padding=(dilation[2], 0, 0), bias=False)
        self.bn4 = norm_layer(planes, momentum=bn_momentum)
        self.conv5 = nn.Conv3d(planes, planes * self.expansion, kernel_size=(1, 1, 1), bias=False)
        self.bn5 = norm_layer(planes * self.expansion, momentum=bn_momentum)
        self.relu = nn.ReLU(inplace=False)
        self.relu_inplace = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.dilation = dilation
        self.stride = stride
        # present (and in every checkpoint) even when stride == 1, reference DDR.py:95-109
        self.downsample2 = nn.Sequential(
            nn.AvgPool3d(kernel_size=(1, stride, 1), stride=(1, stride, 1)),
            nn.Conv3d(planes, planes, kernel_size=1, stride=1, bias=False),
            norm_layer(planes, momentum=bn_momentum))
        self.downsample3 = nn.Sequential(
            nn.AvgPool3d(kernel_size=(stride, 1, 1), stride=(stride, 1, 1)),
            nn.Conv3d(planes, planes, kernel_size=1, stride=1, bias=False),
            norm_layer(planes, momentum=bn_momentum))
        self.downsample4 = nn.Sequential(
            nn.AvgPool3d(kernel_size=(stride, 1, 1), stride=(stride, 1, 1)),
            nn.Conv3d(planes, planes, kernel_size=1, stride=1, bias=False),
            norm_layer(planes, momentum=bn_momentum))

    def emit(self, plan, x, out=None):
        """x: CL -> CL (optionally written into `out`, e.g. a concat window)."""
        s, d = self.stride, self.dilation
        nm = "bneck"
        w, b = fold_bn(self.conv1.weight, None, self.bn1)
        out1 = plan.conv(x, w, b, act="relu", name=nm + ".conv1")
        w, b = fold_bn(self.conv2.weight, None, self.bn2)
        B, D, H, W = out1.dims
        from ..engine import out_size
        o2 = plan.alloc(B, D, H, out_size(W, 3, s, d[0], d[0]), w.shape[0])      # pre-activation out2
        out2r = plan.conv(out1, w, b, stride=(1, 1, s), padding=(0, 0, d[0]), dilation=(1, 1, d[0]), act="relu",
                          out1=o2, out1_mode="cl", name=nm + ".conv2")
        if s != 1:
            wd, bd = fold_bn(self.downsample2[1].weight, None, self.downsample2[2])
            o2 = plan.conv(o2, _pool_conv_weight(wd, (1, s, 1)), bd, stride=(1, s, 1), name=nm + ".ds2")
        w, b = fold_bn(self.conv3.weight, None, self.bn3)
        B, D, H, W = out2r.dims
        o3 = plan.alloc(B, D, out_size(H, 3, s, d[1], d[1]), W, w.shape[0])      # pre-activation out3 (+out2)
        out3r = plan.conv(out2r, w, b, stride=(1, s, 1), padding=(0, d[1], 0), dilation=(1, d[1], 1), act="relu",
                          res1=o2, out1=o3, out1_mode="cl", name=nm + ".conv3")
        if s != 1:
            wd, bd = fold_bn(self.downsample3[1].weight, None, self.downsample3[2])
            o2 = plan.conv(o2, _pool_conv_weight(wd, (s, 1, 1)), bd, stride=(s, 1, 1), name=nm + ".ds3")
            wd, bd = fold_bn(self.downsample4[1].weight, None, self.downsample4[2])
            o3 = plan.conv(o3, _pool_conv_weight(wd, (s, 1, 1)), bd, stride=(s, 1, 1), name=nm + ".ds4")
        w, b = fold_bn(self.conv4.weight, None, self.bn4)
        out4r = plan.conv(out3r, w, b, stride=(s, 1, 1), padding=(d[2], 0, 0), dilation=(d[2], 1, 1), act="relu",
                          res1=o2, res2=o3, name=nm + ".conv4")
        residual = x
        if self.downsample is not None:
            # reference use (modules.py:329-339): AvgPool3d(2,2) -> Conv3d 1x1x1 -> norm
            pool, conv, bn = self.downsample[0], self.downsample[1], self.downsample[2]
            k = pool.kernel_size if isinstance(pool.kernel_size, tuple) else (pool.kernel_size,) * 3
            wd, bd = fold_bn(conv.weight, conv.bias, bn)
            residual = plan.conv(x, _pool_conv_weight(wd, k), bd, stride=k, name=nm + ".ds")
        w, b = fold_bn(self.conv5.weight, None, self.bn5)
        return plan.conv(out4r, w, b, act="relu", res1=residual, out=out, name=nm + ".conv5")

    def forward(self, x):
        return self._run_planar(x)
