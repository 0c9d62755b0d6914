"""A/B the opt-in kernel variants on one B200: each configuration runs in its own process (several switches are
read once per process), times the config-2 forward exactly like bench.py's device-resident arm and compares its
ssc_logit with the default configuration's (relative error + arg-max agreement).

    python tools/ab_experiments.py                  # every single switch + all together
    python tools/ab_experiments.py halox tcx+halox  # chosen combinations

Prints one line per configuration and writes gpurun_out/ab_experiments.json.  AB_PROFILE=1 adds a per-op-group
CUDA-event profile (Plan.profile) of every configuration to the JSON.

Kernel-level parity of the variants first (each in its own process, the C-side switches are read once):
    OCCD_EXPERIMENTAL=1 python -m pytest tests/test_gpu_conv.py tests/test_gpu_ops.py -q -m gpu
    OCCD_EPI_WIDE=1 python -m pytest tests/test_gpu_conv.py -q -m gpu
    OCCD_PDL=1 python -m pytest tests/test_gpu_conv.py tests/test_gpu_unet3d.py -q -m gpu
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SWITCHES = {
    "halox": {"OCCDEPTH_HALOX": "1"},        # x-packed halo kernel (head convs)
    "tcx": {"OCCDEPTH_TCX": "1"},            # x-packed per-tap kernel (Cout <= 80 convs with W taps)
    "tcm2": {"OCCDEPTH_TCM2": "1"},          # two M tiles per weight tile for the wide decoder convs
    "pdl": {"OCCD_PDL": "1"},                # programmatic dependent launch for the conv kernels
    "sestrip": {"OCCDEPTH_SE_IMPL": "strip"},  # SE gate fold, one CTA per 32-channel strip
    "uprows": {"OCCDEPTH_UPSAMPLE_IMPL": "rows"},  # bilinear resize, one block row per output row
    "epiwide": {"OCCD_EPI_WIDE": "1"},        # 256-bit epilogue loads/stores (per-tap kernel, aligned windows)
    "stageslegacy": {"OCCD_TC_STAGES_LEGACY": "1"},  # old ring depth rule (2 x groups per tile) for reference
    "stages16": {"OCCD_TC_STAGES_MIN": "16"},
    "dwdirect": {"OCCDEPTH_DW_IMPL": "direct"},  # the old register-window depthwise kernel (for reference)
}

WORKER = r"""
import json, sys, time, torch
sys.path.insert(0, %(root)r)
import bench
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
m = bench.build_model().to(dev)
img, pix, fov = bench.make_inputs(0)
b = {"img": img.to(dev), "projected_pix_2": [pix.to(dev)], "fov_mask_2": [fov.to(dev)]}
with torch.no_grad():
    for _ in range(4):
        out = m(b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(%(steps)d):
        out = m(b)
    e1.record()
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / %(steps)d
torch.save(out["ssc_logit"].float().cpu(), %(dump)r)
groups = {}
if %(profile)d:
    import re
    plan = list(m._plans().values())[0][0]
    plan.profile()
    for n, t, f in plan.profile():
        k = re.sub(r"\d+", "#", n)
        groups[k] = round(groups.get(k, 0.0) + t, 4)
print("AB_RESULT " + json.dumps({"ms": ms, "profile_ms": groups}))
"""


def run(name, env_extra, steps, dump):
    env = dict(os.environ)
    env.update(env_extra)
    code = WORKER % {"root": ROOT, "steps": steps, "dump": dump, "profile": 1 if os.environ.get("AB_PROFILE") == "1" else 0}
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    for line in r.stdout.splitlines():
        if line.startswith("AB_RESULT "):
            return json.loads(line[len("AB_RESULT "):]), None
    return None, (r.stderr or r.stdout)[-1500:]


def main():
    import torch
    names = sys.argv[1:] or list(SWITCHES) + ["halox+tcx+tcm2+pdl+sestrip+epiwide+uprows"]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    base_dump = "/tmp/ab_default.pt"
    base, err = run("default", {}, 10, base_dump)
    if base is None:
        print("default configuration failed:\n" + err)
        return 1
    ref = torch.load(base_dump)
    results = {"default": base}
    print("%-28s %8.3f ms" % ("default", base["ms"]))
    for name in names:
        env = {}
        for part in name.split("+"):
            env.update(SWITCHES[part])
        dump = "/tmp/ab_%s.pt" % name.replace("+", "_")
        res, err = run(name, env, 10, dump)
        if res is None:
            print("%-28s FAILED: %s" % (name, err.strip().splitlines()[-1] if err.strip() else "?"))
            results[name] = {"error": err}
            continue
        got = torch.load(dump)
        rel = float((got - ref).abs().max() / ref.abs().max())
        agree = float((got.argmax(1) == ref.argmax(1)).float().mean())
        res.update(rel_err_vs_default=rel, argmax_agreement=agree, speedup=base["ms"] / res["ms"])
        results[name] = res
        print("%-28s %8.3f ms  x%.3f  rel %.2e  argmax %.4f" % (name, res["ms"], res["speedup"], rel, agree))
    with open(os.path.join(ROOT, "gpurun_out", "ab_experiments.json"), "w") as f:
        json.dump(results, f, indent=1)
    return 0


if __name__ == "__main__":
    sys.exit(main())
