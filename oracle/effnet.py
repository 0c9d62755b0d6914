"""geffnet-shaped tf_efficientnet_b{3,4,5,7}_ns (oracle side, CPU fp32).

The reference obtains its 2D backbone with
    torch.hub.load("rwightman/gen-efficientnet-pytorch", "tf_efficientnet_b7_ns", pretrained=True)
(occdepth/models/unet2d.py:238-240): un-vendored, unpinned, unreachable offline.  This file restates the
published architecture with geffnet's module names / registration order, which is what the reference's
`Encoder.forward` (unet2d.py:188-196) and released checkpoints (`net_rgb.encoder.original_model.*`) rely on:

  conv_stem, bn1, act1, blocks[0..6], conv_head, bn2, act2, global_pool, classifier

  * width / depth multipliers: b3 (1.2, 1.4), b4 (1.4, 1.8), b5 (1.6, 2.2), b7 (2.0, 3.1)
  * base arch: ds_r1_k3_s1_e1_c16, ir_r2_k3_s2_e6_c24, ir_r2_k5_s2_e6_c40, ir_r3_k3_s2_e6_c80,
               ir_r3_k5_s1_e6_c112, ir_r4_k5_s2_e6_c192, ir_r1_k3_s1_e6_c320; SE ratio 0.25 of block input
  * stem 32*w, head 1280*w channels; Swish (SiLU); BatchNorm eps 1e-3; TensorFlow "SAME" padding
    (asymmetric for stride 2); residual when stride == 1 and in == out; drop-path is identity in eval
Cross-checks available offline: MODEL_CHANNELS / NUM_FEATURES of unet2d.py:10-21, and torchvision's
independent EfficientNet implementation (tests/test_oracle_effnet.py).
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

ARCH = [  # (type, repeats, kernel, stride, expand, out)
    ("ds", 1, 3, 1, 1, 16),
    ("ir", 2, 3, 2, 6, 24),
    ("ir", 2, 5, 2, 6, 40),
    ("ir", 3, 3, 2, 6, 80),
    ("ir", 3, 5, 1, 6, 112),
    ("ir", 4, 5, 2, 6, 192),
    ("ir", 1, 3, 1, 6, 320),
]
MULTS = {
    "tf_efficientnet_b3_ns": (1.2, 1.4),
    "tf_efficientnet_b4_ns": (1.4, 1.8),
    "tf_efficientnet_b5_ns": (1.6, 2.2),
    "tf_efficientnet_b7_ns": (2.0, 3.1),
}
BN_EPS = 1e-3


def round_channels(c, mult, divisor=8):
    c = c * mult
    new_c = max(divisor, int(c + divisor / 2) // divisor * divisor)
    if new_c < 0.9 * c:
        new_c += divisor
    return new_c


def same_pad(i, k, s, d=1):
    return max((math.ceil(i / s) - 1) * s + (k - 1) * d + 1 - i, 0)


class Conv2dSame(nn.Conv2d):
    """TF 'SAME' convolution: pad so out = ceil(in / stride), extra pixel on the bottom/right."""

    def forward(self, x):
        ih, iw = x.shape[-2:]
        kh, kw = self.weight.shape[-2:]
        ph = same_pad(ih, kh, self.stride[0], self.dilation[0])
        pw = same_pad(iw, kw, self.stride[1], self.dilation[1])
        if ph > 0 or pw > 0:
            x = F.pad(x, [pw // 2, pw - pw // 2, ph // 2, ph - ph // 2])
        return F.conv2d(x, self.weight, self.bias, self.stride, (0, 0), self.dilation, self.groups)


def conv2d_tf(cin, cout, k, stride=1, groups=1, bias=False):
    if stride == 1:
        return nn.Conv2d(cin, cout, k, 1, (k - 1) // 2, groups=groups, bias=bias)
    return Conv2dSame(cin, cout, k, stride, 0, groups=groups, bias=bias)


class SqueezeExcite(nn.Module):
    def __init__(self, chs, reduce_chs):
        super().__init__()
        self.conv_reduce = nn.Conv2d(chs, reduce_chs, 1, bias=True)
        self.act1 = nn.SiLU()
        self.conv_expand = nn.Conv2d(reduce_chs, chs, 1, bias=True)

    def forward(self, x):
        s = x.mean((2, 3), keepdim=True)
        s = self.conv_expand(self.act1(self.conv_reduce(s)))
        return x * torch.sigmoid(s)


class DepthwiseSeparableConv(nn.Module):
    def __init__(self, cin, cout, k, stride, se_chs):
        super().__init__()
        self.has_residual = stride == 1 and cin == cout
        self.conv_dw = conv2d_tf(cin, cin, k, stride, groups=cin)
        self.bn1 = nn.BatchNorm2d(cin, eps=BN_EPS)
        self.act1 = nn.SiLU()
        self.se = SqueezeExcite(cin, se_chs)
        self.conv_pw = nn.Conv2d(cin, cout, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(cout, eps=BN_EPS)
        self.act2 = nn.Identity()

    def forward(self, x):
        y = self.act1(self.bn1(self.conv_dw(x)))
        y = self.se(y)
        y = self.bn2(self.conv_pw(y))
        return y + x if self.has_residual else y


class InvertedResidual(nn.Module):
    def __init__(self, cin, cout, k, stride, expand, se_chs):
        super().__init__()
        mid = cin * expand
        self.has_residual = stride == 1 and cin == cout
        self.conv_pw = nn.Conv2d(cin, mid, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(mid, eps=BN_EPS)
        self.act1 = nn.SiLU()
        self.conv_dw = conv2d_tf(mid, mid, k, stride, groups=mid)
        self.bn2 = nn.BatchNorm2d(mid, eps=BN_EPS)
        self.act2 = nn.SiLU()
        self.se = SqueezeExcite(mid, se_chs)
        self.conv_pwl = nn.Conv2d(mid, cout, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(cout, eps=BN_EPS)

    def forward(self, x):
        y = self.act1(self.bn1(self.conv_pw(x)))
        y = self.act2(self.bn2(self.conv_dw(y)))
        y = self.se(y)
        y = self.bn3(self.conv_pwl(y))
        return y + x if self.has_residual else y


def block_specs(name):
    """[(stage, type, cin, cout, k, stride, expand, se_chs), ...] for every block, plus stem/head chs."""
    wm, dm = MULTS[name]
    stem = round_channels(32, wm)
    specs = []
    cin = stem
    for si, (typ, r, k, s, e, c) in enumerate(ARCH):
        cout = round_channels(c, wm)
        reps = int(math.ceil(r * dm))
        for bi in range(reps):
            stride = s if bi == 0 else 1
            specs.append((si, typ, cin, cout, k, stride, e, max(1, int(cin * 0.25 + 0.5))))
            cin = cout
    head = round_channels(1280, wm)
    return stem, specs, head


class GenEfficientNet(nn.Module):
    def __init__(self, name="tf_efficientnet_b7_ns", num_classes=1000):
        super().__init__()
        stem, specs, head = block_specs(name)
        self.conv_stem = conv2d_tf(3, stem, 3, 2)
        self.bn1 = nn.BatchNorm2d(stem, eps=BN_EPS)
        self.act1 = nn.SiLU()
        stages = [[] for _ in ARCH]
        for (si, typ, cin, cout, k, stride, e, se) in specs:
            if typ == "ds":
                stages[si].append(DepthwiseSeparableConv(cin, cout, k, stride, se))
            else:
                stages[si].append(InvertedResidual(cin, cout, k, stride, e, se))
        self.blocks = nn.Sequential(*[nn.Sequential(*s) for s in stages])
        self.conv_head = nn.Conv2d(specs[-1][3], head, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(head, eps=BN_EPS)
        self.act2 = nn.SiLU()
        self.global_pool = nn.AdaptiveAvgPool2d(1)
        self.classifier = nn.Linear(head, num_classes)

    def forward(self, x):
        x = self.act1(self.bn1(self.conv_stem(x)))
        x = self.blocks(x)
        x = self.act2(self.bn2(self.conv_head(x)))
        return self.classifier(self.global_pool(x).flatten(1))
