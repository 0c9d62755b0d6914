"""Pins oracle/effnet.py (geffnet-shaped EfficientNet restatement) against torchvision's INDEPENDENT EfficientNet
implementation: same weights, an input whose every stride-2 stage sees an odd size (65 -> 33 -> 17 -> 9 -> 5 -> 3) so
that TensorFlow "SAME" padding coincides with torchvision's symmetric padding.  Also checks the channel tables the
reference hard-codes (unet2d.py:10-21)."""
import pytest
import torch
import torch.nn as nn

from oracle import effnet

tv = pytest.importorskip("torchvision.models")


def _copy_conv_bn(dst_conv, dst_bn, src):
    dst_conv.weight.data.copy_(src[0].weight.data)
    for k in ("weight", "bias", "running_mean", "running_var"):
        getattr(dst_bn, k).data.copy_(getattr(src[1], k).data)


def test_effnet_b3_matches_torchvision():
    torch.manual_seed(0)
    ref = tv.efficientnet_b3(weights=None).eval()
    for m in ref.modules():
        if isinstance(m, nn.BatchNorm2d):
            m.eps = effnet.BN_EPS
            m.running_mean.normal_(0, 0.1)
            m.running_var.uniform_(0.5, 1.5)
            m.weight.data.uniform_(0.5, 1.5)
            m.bias.data.normal_(0, 0.1)
    mine = effnet.GenEfficientNet("tf_efficientnet_b3_ns").eval()
    _copy_conv_bn(mine.conv_stem, mine.bn1, ref.features[0])
    for si in range(7):
        assert len(ref.features[si + 1]) == len(mine.blocks[si]), si
        for bi, blk in enumerate(mine.blocks[si]):
            parts = list(ref.features[si + 1][bi].block.children())
            if isinstance(blk, effnet.DepthwiseSeparableConv):
                dw, se, proj = parts
                _copy_conv_bn(blk.conv_dw, blk.bn1, dw)
                _copy_conv_bn(blk.conv_pw, blk.bn2, proj)
            else:
                ex, dw, se, proj = parts
                _copy_conv_bn(blk.conv_pw, blk.bn1, ex)
                _copy_conv_bn(blk.conv_dw, blk.bn2, dw)
                _copy_conv_bn(blk.conv_pwl, blk.bn3, proj)
            assert blk.se.conv_reduce.weight.shape == se.fc1.weight.shape, (si, bi)
            blk.se.conv_reduce.load_state_dict(se.fc1.state_dict())
            blk.se.conv_expand.load_state_dict(se.fc2.state_dict())
    _copy_conv_bn(mine.conv_head, mine.bn2, ref.features[8])
    x = torch.randn(1, 3, 65, 65)
    with torch.no_grad():
        want = ref.features(x)
        got = mine.act2(mine.bn2(mine.conv_head(mine.blocks(mine.act1(mine.bn1(mine.conv_stem(x)))))))
    assert got.shape == want.shape == (1, 1536, 3, 3)
    assert torch.allclose(got, want, rtol=1e-4, atol=1e-5), float((got - want).abs().max())


def test_channel_tables_match_reference_constants():
    # occdepth/models/unet2d.py:10-21 (MODEL_CHANNELS = [image, blocks[0], blocks[1], blocks[2], blocks[4]], NUM_FEATURES)
    expect = {"tf_efficientnet_b3_ns": ([24, 32, 48, 136], 1536), "tf_efficientnet_b4_ns": ([24, 32, 56, 160], 1792),
              "tf_efficientnet_b5_ns": ([24, 40, 64, 176], 2048), "tf_efficientnet_b7_ns": ([32, 48, 80, 224], 2560)}
    for name, (chs, head) in expect.items():
        stem, specs, h = effnet.block_specs(name)
        outs = [[s for s in specs if s[0] == i][-1][3] for i in range(7)]
        assert [outs[0], outs[1], outs[2], outs[4]] == chs and h == head, name
    m = effnet.GenEfficientNet("tf_efficientnet_b7_ns")
    assert abs(sum(p.numel() for p in m.parameters()) / 1e6 - 66.35) < 0.1     # published B7 size
