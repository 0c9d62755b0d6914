"""Standalone depthwise-conv timing (B=2 views): python tools/dw_bench.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from occdepth_b200 import _lib
L = _lib.lib()
dev = torch.device("cuda")
PREC = os.environ.get("OCCDEPTH_PRECISION", "tf32")
TDT, CODE, ES = (torch.float32, 0, 4) if PREC == "tf32" else (torch.bfloat16, 1, 2)
FN = L.occd_dwconv2d_fwd if os.environ.get("DW_IMPL") == "direct" else L.occd_dwconv2d_tiled_fwd
SHAPES = {"b0_c64_k3": (64, 188, 685, 3, 1), "b1_c288_k3": (288, 94, 343, 3, 1), "b1_c192_k3s2": (192, 188, 685, 3, 2),
          "b2_c480_k5": (480, 47, 172, 5, 1), "b4_c1344_k5": (1344, 24, 86, 5, 1), "b5_c2304_k5": (2304, 12, 43, 5, 1)}
names = sys.argv[1:] or list(SHAPES)
for n in names:
    C, H, W, K, S = SHAPES[n]
    B = 2
    OH, OW = (H + S - 1) // S, (W + S - 1) // S
    x = torch.randn(B, H, W, C, device=dev).to(TDT)
    y = torch.empty(B, OH, OW, C, device=dev, dtype=TDT)
    w = torch.randn(K * K, C, device=dev)
    b = torch.randn(C, device=dev)
    pool = torch.zeros(B, C, dtype=torch.int64, device=dev)
    pad = max((OH - 1) * S + K - H, 0) // 2
    for use_pool in (True, False):
        def run():
            rc = FN(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(),
                    pool.data_ptr() if use_pool else None, CODE, B, H, W, OH, OW, C, C, C, K, S, pad, pad,
                    _lib.ACT_SILU, _lib.stream_ptr())
            assert rc == 0
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        byts = (x.numel() + y.numel()) * ES
        print("%-14s pool=%d %8.4f ms  %7.1f GB/s" % (n, use_pool, ms, byts / ms / 1e6), flush=True)
