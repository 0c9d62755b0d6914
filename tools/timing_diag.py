"""Why did the first timed region of bench.py read 34 ms when every later one read ~20 ms?  Repeat the device-resident
timed loop several times in one process and print each (plus SM clocks from nvidia-smi around them)."""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402


def clocks():
    o = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.active",
                        "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
    return o


dev = torch.device("cuda", 0)
m = bench.build_model().to(dev).set_precision(sys.argv[1] if len(sys.argv) > 1 else "tf32")
img, pix, fov = bench.make_inputs(0)
b = {"img": img.to(dev), "projected_pix_2": [pix.to(dev)], "fov_mask_2": [fov.to(dev)]}
with torch.no_grad():
    t0 = time.time()
    for _ in range(3):
        m(b)
    torch.cuda.synchronize()
    print("3 warm-up forwards (incl. plan build): %.2f s" % (time.time() - t0), clocks(), flush=True)
    for rep in range(8):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0 = time.perf_counter()
        e0.record()
        for _ in range(10):
            m(b)
        e1.record()
        c_enq = time.perf_counter() - c0
        torch.cuda.synchronize()
        print("rep %d: %.3f ms/step (host enqueue %.3f ms/step)  %s" % (rep, e0.elapsed_time(e1) / 10, c_enq * 100,
                                                                      clocks()), flush=True)
    # the same with the graph replay alone (no input staging, no output conversion)
    plan = list(m._plans().values())[0][0]
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            plan.run()
        e1.record()
        torch.cuda.synchronize()
        print("plan.run only rep %d: %.3f ms/step" % (rep, e0.elapsed_time(e1) / 10), flush=True)
