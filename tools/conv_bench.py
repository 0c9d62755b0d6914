"""Standalone timing of representative conv shapes (for ncu and quick A/B): python tools/conv_bench.py [names...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from occdepth_b200.engine import CL, ConvOp, Plan, conv_taps  # noqa: E402

SHAPES = {
    # name: (dims (D,H,W), Cin, Cout, kernel, dilation, act)
    "expand_188": ((1, 188, 685), 32, 192, (1, 1, 1), 1, "silu"),
    "proj_94": ((1, 94, 343), 288, 48, (1, 1, 1), 1, "none"),
    "expand_24": ((1, 24, 86), 224, 1344, (1, 1, 1), 1, "silu"),
    "head_c32_d1": ((256, 256, 32), 32, 32, (3, 3, 3), 1, "relu"),
    "head_c32_d3": ((256, 256, 32), 32, 32, (3, 3, 3), 3, "relu"),
    "up1_conv2": ((1, 376, 1370), 80, 80, (1, 3, 3), 1, "leaky"),
    "up2_conv2": ((1, 188, 685), 160, 160, (1, 3, 3), 1, "leaky"),
    "up16_conv1": ((1, 24, 86), 2784, 1280, (1, 3, 3), 1, "leaky"),
    "head_c32_d2": ((256, 256, 32), 32, 32, (3, 3, 3), 2, "relu"),
    "head_c32_n2": ((256, 256, 32), 32, 2, (3, 3, 3), 1, "none"),
    "b5_expand": ((1, 24, 43), 384, 2304, (1, 1, 1), 1, "silu"),
    "b5_proj": ((1, 24, 43), 2304, 384, (1, 1, 1), 1, "none"),
    "x_c32_n32": ((256, 256, 16), 32, 32, (3, 3, 3), 1, "relu"),
    "x_c64_n32": ((256, 256, 16), 64, 32, (3, 3, 3), 1, "relu"),
    "x_c32_n64": ((256, 256, 16), 32, 64, (3, 3, 3), 1, "relu"),
    "x_c64_n64": ((256, 256, 16), 64, 64, (3, 3, 3), 1, "relu"),
    "x_c64_n256": ((128, 128, 16), 64, 256, (3, 3, 3), 1, "relu"),
    "x_c16_n16": ((256, 256, 16), 16, 16, (3, 3, 3), 1, "relu"),
    "b1_expand": ((1, 94, 686), 48, 288, (1, 1, 1), 1, "silu"),      # both views side by side (B = 2 in the net)
    "b1_expand_noact": ((1, 94, 686), 48, 288, (1, 1, 1), 1, "none"),
    "b1_expand_relu": ((1, 94, 686), 48, 288, (1, 1, 1), 1, "relu"),
    "bneck5_16_64": ((128, 128, 16), 16, 64, (1, 1, 1), 1, "relu"),
    "up4_conv2": ((1, 94, 686), 320, 320, (1, 3, 3), 1, "leaky"),
    "l1_k1_64_16": ((128, 128, 16), 64, 16, (1, 1, 1), 1, "relu"),
    "l1_k113_16": ((128, 128, 16), 16, 16, (1, 1, 3), 1, "relu"),
}


def main():
    names = sys.argv[1:] or list(SHAPES)
    dev = torch.device("cuda")
    for n in names:
        dims, ci, co, k, dl, act = SHAPES[n]
        plan = Plan(dev, precision=os.environ.get("OCCDEPTH_PRECISION", "tf32"))
        x = plan.alloc(1, dims[0], dims[1], dims[2], ci)
        x.buf.normal_()
        w = torch.randn(co, ci, *k, device=dev) / (ci * k[0] * k[1] * k[2]) ** 0.5
        b = torch.randn(co, device=dev)
        pad = tuple(dl * (kk - 1) // 2 for kk in k)
        impl = {"tc": 0, "simt": 1, "halo": 2, "tcx": 4}.get(os.environ.get("BENCH_IMPL", ""), None)
        plan.conv(x, w, b, padding=pad, dilation=dl, act=act, name=n, impl=impl)
        for _ in range(3):
            plan.run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        reps = 10
        for _ in range(reps):
            plan.run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        fl = plan.ops[0].flops
        byts = (x.buf.numel() + dims[0] * dims[1] * dims[2] * co) * x.esize
        print("%-14s %8.3f ms  %7.1f TF/s  %7.1f GB/s(min traffic)  %s" % (n, ms, fl / ms / 1e9, byts / ms / 1e6,
                                                                          plan.ops[0].info()), flush=True)


if __name__ == "__main__":
    main()
