#!/usr/bin/env python
"""Benchmark of the OccDepth forward hot path (BASELINE.json metric: forward voxels/sec).

  python bench.py --gpus N --steps K --warmup W            # B200 arm (one process per GPU under torchrun)
  python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host CPU cores

Workload (BASELINE.json configs[1]): synthetic 1370x376 stereo pair, tf_efficientnet_b7_ns 2D backbone,
Stereo-SFA lift to 128x128x16 / 64 ch, 3D UNet + CRP + cascade head -> 256x256x32 voxel logits (20 classes).
A step = one OccDepth.forward over one frame per GPU (frames are independent: replicas, weak scaling).
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

IMG_H, IMG_W = 376, 1370
FULL = (256, 256, 32)
N_OUT = FULL[0] * FULL[1] * FULL[2]
PROJECT_RES = ["1", "2", "4", "8"]
WORKLOAD = ("configs[1]: 1370x376 stereo, tf_efficientnet_b7_ns, flosp lift to 128x128x16/64ch, "
            "UNet3D+CRP+cascade head -> 256x256x32x20 logits, B=1 per GPU")


def make_cfg():
    import synthetic as synth
    return synth.occdepth_cfg(full_scene_size=FULL, project_scale=2, feature=64, feature_2d_oc=64, n_classes=20,
                              backbone_2d_name="tf_efficientnet_b7_ns", cascade_cls=True, context_prior=True)


def make_inputs(seed=0):
    import torch
    import synthetic as synth
    g = torch.Generator().manual_seed(seed)
    img = torch.randn(1, 2, 3, IMG_H, IMG_W, generator=g)
    pix, fov, _, _ = synth.kitti_indices(IMG_W, IMG_H, FULL, 2, voxel=0.2)
    return img, pix, fov


def build_model():
    import contextlib
    import io
    import torch
    import synthetic as synth
    from occdepth_b200.models.OccDepth import OccDepth
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        m = OccDepth(["c"] * 20, torch.ones(20), full_scene_size=FULL, project_res=PROJECT_RES, config=make_cfg())
    synth.randomize_bn_(m)
    return m.eval()


class ClockSampler(threading.Thread):
    """samples nvidia-smi clocks / throttle reasons while the timed region runs"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                    "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in o.strip().split(",")]
                if len(f) >= 6:
                    self.samples.append(f)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(float(s[0])) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": int(float(self.samples[0][1])), "reasons": reasons,
                "samples": len(sm)}


def cpu_forward_seconds(steps=1, warmup=0, threads=None):
    """the reference algorithm (oracle/functional.py, pinned against /root/reference) on the host cores"""
    import torch
    from oracle import functional as OF      # the ONLY leg of bench.py that executes oracle/ code
    if threads:
        torch.set_num_threads(threads)
    m = build_model()
    sd = {k: v for k, v in m.state_dict().items()}
    img, pix, fov = make_inputs()
    batch = {"img": img, "projected_pix_2": [pix], "fov_mask_2": [fov]}
    cfg = dict(make_cfg())
    cfg["project_res"] = PROJECT_RES
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            OF.occdepth_forward(sd, batch, cfg)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    return times, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    times, cores = cpu_forward_seconds(args.steps, args.warmup)
    total = sum(times)
    v = N_OUT * len(times) / total
    line = {
        "impl": "reference", "metric": "forward voxels/sec", "value": v, "unit": "voxels/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD},
        "cpu_baseline": {"value": v, "unit": "voxels/s", "cores": cores, "kind": "port",
                         "sample": "%d full forward(s) of the workload (oracle/functional.py, PyTorch CPU fp32)" % len(times)},
        "e2e": {"value": v, "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def run_b200(args):
    import torch
    import torch.distributed as dist
    from occdepth_b200 import parallel
    world, rank, local = parallel.env_world()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    parallel.init("nccl", dev)
    m = build_model().to(dev)
    slab = args.partition == "slab" and world > 1
    img, pix, fov = make_inputs(seed=0 if slab else rank)
    if slab:
        # BASELINE.json configs[2]: ONE frame, voxel grid split along X over the ranks, NCCL halo exchange at the
        # 3-D conv boundaries (strong scaling); the default is one independent frame per rank (weak scaling)
        m.enable_slab_parallel(parallel.SlabContext(halo=3))
    # ---- device-resident arm ----
    batch_dev = {"img": img.to(dev), "projected_pix_2": [pix.to(dev)], "fov_mask_2": [fov.to(dev)]}
    with torch.no_grad():
        for _ in range(max(args.warmup, 3)):
            out = m(batch_dev)
    torch.cuda.synchronize()

    def barrier():
        parallel.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    with torch.no_grad():
        for _ in range(args.steps):
            out = m(batch_dev)
    e1.record()
    barrier()
    ms_total = parallel.max_over_ranks(e0.elapsed_time(e1), dev)

    # ---- end-to-end arm: host buffers, H2D of the inputs and D2H of the logits inside the timed region ----
    img_h, pix_h, fov_h = img.pin_memory(), pix.pin_memory(), fov.pin_memory()
    logits_h = torch.empty(out["ssc_logit"].shape, dtype=torch.float32).pin_memory()
    h2d = img_h.numel() * 4 + pix_h.numel() * 8 + fov_h.numel()
    d2h = logits_h.numel() * 4

    def e2e_step():
        b = {"img": img_h.to(dev, non_blocking=True), "projected_pix_2": [pix_h], "fov_mask_2": [fov_h]}
        o = m(b)
        logits_h.copy_(o["ssc_logit"], non_blocking=True)

    with torch.no_grad():
        for _ in range(3):
            e2e_step()
        barrier()
        e0.record()
        for _ in range(args.steps):
            e2e_step()
        e1.record()
    barrier()
    ms_e2e = parallel.max_over_ranks(e0.elapsed_time(e1), dev)

    # ---- end-to-end, pipelined: the same per-step copies, but the D2H read of step i runs on a copy stream into
    # double-buffered pinned host buffers while the forward of step i+1 runs (what a serving loop does).  No
    # collective inside the try block: a rank that fails reports inf and every rank falls back to the sequential
    # number above.
    ms_pipe_local = float("inf")
    pipe_err = None
    if not slab:
        try:
            cs = torch.cuda.Stream(device=dev)
            main = torch.cuda.current_stream(dev)
            host_bufs = [logits_h, torch.empty(logits_h.shape, dtype=torch.float32).pin_memory()]

            def pipe_step(i):
                b = {"img": img_h.to(dev, non_blocking=True), "projected_pix_2": [pix_h], "fov_mask_2": [fov_h]}
                res = m(b)["ssc_logit"]
                cs.wait_stream(main)          # the forward (and its output copy) queued so far
                res.record_stream(cs)
                with torch.cuda.stream(cs):
                    host_bufs[i & 1].copy_(res, non_blocking=True)

            with torch.no_grad():
                for i in range(3):
                    pipe_step(i)
                main.wait_stream(cs)
                torch.cuda.synchronize()
                p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                p0.record()
                for i in range(args.steps):
                    pipe_step(i)
                main.wait_stream(cs)          # the last D2H read ends inside the timed region
                p1.record()
                torch.cuda.synchronize()
            ms_pipe_local = p0.elapsed_time(p1)
        except Exception as ex:  # noqa: BLE001
            pipe_err = repr(ex)
            torch.cuda.synchronize()
    # ---- end-to-end of the widened path: per step the host supplies the image and the
    # calibration only -- the projection indices are generated on the device (occdepth_b200.data.vox2pix, the data
    # pipeline's numba job) and the reference callers' softmax/argmax post-processing (generate_output.py:94-97) runs
    # on the device too (OccDepth.predict), so the D2H read is the 4 MB uint16 class map instead of 168 MB of logits.
    # Reported next to, not instead of, `e2e`.
    ms_cls_local = float("inf")
    cls_err = None
    d2h_cls = 0
    if not slab:
        try:
            import numpy as np
            import synthetic as synth
            from occdepth_b200.data import normalize_rgb as normalize_rgb_dev, vox2pix as vox2pix_dev
            Kc, Tc = synth.kitti_calib(IMG_W, IMG_H)
            # camera frames are uint8: the image crosses PCIe as uint8 and is normalised on the device
            # (occdepth_b200.data.normalize_rgb = the datasets' /255 + ToTensor + Normalize)
            u8_h = torch.randint(0, 256, (2, IMG_H, IMG_W, 3), dtype=torch.uint8,
                                 generator=torch.Generator().manual_seed(7)).pin_memory()
            scene_m = tuple(v * 0.2 for v in FULL)
            origin = np.array([0.0, -scene_m[1] / 2.0, -2.0])

            def indices_on_device():
                pv, fv = [], []
                for T in Tc:
                    p, f, _ = vox2pix_dev(T, Kc, origin, 0.4, IMG_W, IMG_H, scene_m, 0, device=dev)
                    pv.append(p)
                    fv.append(f)
                return torch.stack(pv), torch.stack(fv)

            with torch.no_grad():
                p_dev, f_dev = indices_on_device()
                if not (torch.equal(p_dev.cpu(), pix) and torch.equal(f_dev.cpu(), fov)):
                    raise RuntimeError("device vox2pix differs from the host indices")
                cls_h = None
                for it in range(3 + args.steps):
                    if it == 3:
                        torch.cuda.synchronize()
                        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        c0.record()
                    p_dev, f_dev = indices_on_device()
                    u8_d = u8_h.to(dev, non_blocking=True)
                    img_d = torch.stack([normalize_rgb_dev(u8_d[v], IMG_H, IMG_W, device=dev) for v in range(2)])[None]
                    b = {"img": img_d, "projected_pix_2": [p_dev], "fov_mask_2": [f_dev]}
                    y, _ = m.predict(b)
                    if cls_h is None:
                        cls_h = torch.empty(y.shape, dtype=y.dtype).pin_memory()
                    cls_h.copy_(y, non_blocking=True)
                c1.record()
                torch.cuda.synchronize()
            ms_cls_local = c0.elapsed_time(c1)
            d2h_cls = cls_h.numel() * cls_h.element_size()
        except Exception as ex:  # noqa: BLE001
            cls_err = repr(ex)
            torch.cuda.synchronize()
    ms_cls = parallel.max_over_ranks(ms_cls_local, dev)
    ms_pipe = parallel.max_over_ranks(ms_pipe_local, dev)
    pipelined = ms_pipe != float("inf") and ms_pipe > 0
    ms_e2e_best = min(ms_pipe, ms_e2e) if pipelined else ms_e2e
    # keep the GPU under the same load a little longer so that nvidia-smi (100 ms period) sees it, then stop
    with torch.no_grad():
        if slab:      # every rank must run the SAME number of forwards (each one is a set of NCCL exchanges)
            for _ in range(40):
                m(batch_dev)
        else:
            t_end = time.time() + 1.0
            while time.time() < t_end:
                m(batch_dev)
        torch.cuda.synchronize()
    sampler.stop_flag = True

    line = None
    if rank == 0:
        plan = list(m._plans().values())[0][0]
        # per-kernel profile pass (outside the timed regions): shares + roofline of the dominant kernel
        # (not in slab mode: the plan contains NCCL ops that every rank would have to enter together)
        prof = [] if slab else (plan.profile(), plan.profile())[1]
        conv_ms = sum(t for n, t, f in prof if f > 0)
        conv_fl = sum(f for n, t, f in prof if f > 0)
        lift_ms = sum(t for n, t, f in prof if n == "sfa_lift")
        tot_ms = sum(t for n, t, f in prof)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:  # noqa: BLE001
            pass
        tpeak = peaks.get("bf16_tflops_sustained", 1400.0)
        hpeak = peaks.get("hbm_gbs", 6650.0)
        src = "measured" if peaks else "fallback"
        # lift algorithmic bytes (SURVEY 8d formula with the element sizes actually used: bf16 features / output)
        U = 0
        for s in (1, 2, 4, 8):
            import synthetic as synth
            h, w = synth.feature_hw(IMG_H, IMG_W, s)
            for v in range(2):
                idx = (pix[v, :, 0, 1] // s) * w + (pix[v, :, 0, 0] // s)
                U += int(torch.unique(idx[fov[v, :, 0]]).numel())
        N1 = pix.shape[1]
        lift_bytes = U * 64 * 2 + 2 * N1 * 17 + N1 * 64 * 2
        ach_t = conv_fl / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
        ach_h = lift_bytes / (lift_ms * 1e-3) / 1e9 if lift_ms > 0 else 0.0
        frames = 1 if slab else world
        value = frames * N_OUT * args.steps / (ms_total * 1e-3)
        cpu = None
        if world == 1 and not args.no_cpu:
            times, cores = cpu_forward_seconds(1, 0)
            cpu = {"value": N_OUT / times[0], "unit": "voxels/s", "cores": cores, "kind": "port",
                   "sample": "1 full forward of the workload (oracle/functional.py, PyTorch CPU fp32), %.1f s" % times[0]}
        line = {
            "metric": "forward voxels/sec", "value": value, "unit": "voxels/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "strong" if slab else "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_step": frames,
                       "partition": "x-slab of one frame + NCCL halo exchange" if slab else "frame replicas",
                       "l2": "per-step working set (weights + activations, >2 GB) exceeds the 126 MB L2; no flush",
                       "cuda_graph": os.environ.get("OCCDEPTH_CUDA_GRAPH", "1") == "1"},
            "e2e": {"value": frames * N_OUT * args.steps / (ms_e2e_best * 1e-3), "unit": "voxels/s",
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e_best / args.steps,
                    "mode": ("pipelined: every step copies its inputs H2D and its logits D2H; the D2H read of step i "
                             "(copy stream, double-buffered pinned host buffers) overlaps the forward of step i+1"
                             if pipelined and ms_pipe <= ms_e2e else "sequential: H2D, forward, D2H back to back"),
                    "sequential_ms_per_step": ms_e2e / args.steps,
                    "pipelined_ms_per_step": ms_pipe / args.steps if pipelined else None,
                    "pipelined_error": pipe_err},
            "e2e_classes": ({"value": frames * N_OUT * args.steps / (ms_cls * 1e-3), "unit": "voxels/s",
                             "ms_per_step": ms_cls / args.steps, "h2d_bytes_per_step": 2 * IMG_H * IMG_W * 3,
                             "d2h_bytes_per_step": d2h_cls,
                             "what": "uint8 stereo frame + calibration in, class map out: normalisation on the device "
                                     "(occdepth_b200.data.normalize_rgb = kitti_dataset.py:376-402), projection indices "
                                     "generated on the device (occdepth_b200.data.vox2pix = helpers.py:94-169), forward, arg-max class "
                                     "map on the device (OccDepth.predict = generate_output.py:94-97), uint16 map "
                                     "read back"}
                            if ms_cls != float("inf") and ms_cls > 0 else {"error": cls_err}),
            "gpu_launches": len(plan.ops) * args.steps,
            "clocks": sampler.summary(),
            "roofline": {"kernel": "conv_tc_kernel + conv_halo_kernel (tcgen05 implicit GEMM family, %d launches/step)" % sum(1 for n, t, f in prof if f > 0),
                         "bound": "tensor", "achieved": ach_t, "peak": tpeak, "unit": "TFLOP/s",
                         "frac": ach_t / tpeak if prof else None, "traffic": None, "peak_source": src + " bf16_tflops_sustained",
                         "note": "achieved = sum of algorithmic FLOPs (2*MACs of the reference convs) / sum of the family's "
                                 "launch durations (CUDA events, back-to-back); per-shape ncu traffic: profiles/r01_ncu_summary.md",
                         "share_of_step": conv_ms / tot_ms if tot_ms else None,
                         "algorithmic_flops_per_step": conv_fl},
            "roofline_lift": {"kernel": "sfa_lift_kernel", "bound": "hbm", "achieved": ach_h, "peak": hpeak,
                              "unit": "GB/s", "frac": ach_h / hpeak, "traffic": 76079360,   # ncu r01b_lift: dram read+write
                              "peak_source": src + " hbm_gbs",
                              "algorithmic_bytes": lift_bytes, "ms": lift_ms,
                              "share_of_step": lift_ms / tot_ms if tot_ms else None},
            "profile_ms": {"convs": conv_ms, "lift": lift_ms, "other": tot_ms - conv_ms - lift_ms, "sum": tot_ms},
        }
        if cpu:
            line["cpu_baseline"] = cpu
        if args.dump_profile:
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            with open(os.path.join(ROOT, "gpurun_out", "plan_profile.json"), "w") as f:
                json.dump([{"name": n, "ms": t, "flops": fl} for n, t, fl in prof], f)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line:
        print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--partition", default="frames", choices=["frames", "slab"],
                    help="frames: one frame per GPU (default, weak scaling); slab: one frame, X-slab partition")
    ap.add_argument("--dump-profile", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    run_b200(args)


if __name__ == "__main__":
    main()
