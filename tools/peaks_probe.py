"""One-off probes on a B200 (results land in gpurun_out/peaks_probe.json):
  * cuBLAS matmul peaks for the roofline denominators the driver does not measure: TF32 (allow_tf32) and plain
    fp32, burst (best of 10) and sustained (back to back for ~3 s), measured the way MEASURED_PEAKS.json measures bf16
  * the reference algorithm (oracle/functional.py, pinned against /root/reference) under eager CUDA on this GPU:
    cuDNN / ATen library kernels, TF32 allowed and not allowed -- the "honest competitor" of SURVEY 8(d)
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def matmul_peak(dtype, allow_tf32, n=8192, seconds=3.0):
    torch.backends.cuda.matmul.allow_tf32 = allow_tf32
    a = torch.randn(n, n, device="cuda", dtype=dtype)
    b = torch.randn(n, n, device="cuda", dtype=dtype)
    for _ in range(3):
        a @ b
    torch.cuda.synchronize()
    best = 0.0
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        a @ b
        e1.record()
        torch.cuda.synchronize()
        best = max(best, 2 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_end = time.time() + seconds
    iters = 0
    e0.record()
    while time.time() < t_end:
        for _ in range(10):
            a @ b
        iters += 10
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    sustained = iters * 2 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12
    return {"burst_tflops": best, "sustained_tflops": sustained}


def eager_cuda_reference(allow_tf32, steps=3):
    import bench
    from oracle import functional as OF
    torch.backends.cuda.matmul.allow_tf32 = allow_tf32
    torch.backends.cudnn.allow_tf32 = allow_tf32
    torch.backends.cudnn.benchmark = True
    dev = torch.device("cuda")
    m = bench.build_model()
    sd = {k: v.to(dev) for k, v in m.state_dict().items()}
    img, pix, fov = bench.make_inputs()
    batch = {"img": img.to(dev), "projected_pix_2": [pix.to(dev)], "fov_mask_2": [fov.to(dev)]}
    cfg = dict(bench.make_cfg())
    cfg["project_res"] = bench.PROJECT_RES
    with torch.no_grad():
        for _ in range(2):
            out = OF.occdepth_forward(sd, batch, cfg)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = OF.occdepth_forward(sd, batch, cfg)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"ms_per_frame": ms, "voxels_per_s": bench.N_OUT / (ms * 1e-3), "allow_tf32": allow_tf32,
            "cudnn_benchmark": True, "logit_absmax": float(out["ssc_logit"].abs().max())}


def main():
    res = {"gpu": torch.cuda.get_device_name(0)}
    res["tf32_matmul"] = matmul_peak(torch.float32, True)
    res["fp32_matmul"] = matmul_peak(torch.float32, False, n=4096, seconds=1.5)
    res["bf16_matmul"] = matmul_peak(torch.bfloat16, True)
    for tf32 in (True, False):
        try:
            res["eager_cuda_tf32" if tf32 else "eager_cuda_fp32"] = eager_cuda_reference(tf32)
        except Exception as ex:  # noqa: BLE001
            res["eager_cuda_tf32" if tf32 else "eager_cuda_fp32"] = {"error": repr(ex)}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "peaks_probe.json"), "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
