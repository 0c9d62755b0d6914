"""Bring-up diagnostics on a real B200: every case in its own subprocess (a trap/hang poisons the context).
usage: python tools/gpu_diag.py [sfa|simt|tc|all] -> gpurun_out/diag.txt"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def run_one(kind, name):
    import torch
    import gpu_cases as G
    from occdepth_b200 import _lib
    impl = {"tc": _lib.CONV_IMPL_TC, "simt": _lib.CONV_IMPL_SIMT, "halo": _lib.CONV_IMPL_HALO}[kind]
    if kind == "halo":
        e, info = G.conv_case(impl, **G.HALO_CASES[name])
    elif name == "convT":
        e, info = G.convT_case(impl)
    elif name == "multi":
        e, info = G.multi_case(impl)
    else:
        e, info = G.conv_case(impl, **G.CONV_CASES[name])
    print(json.dumps(dict(kind=kind, case=name, rel_err=e, info=info)))


def run_many(kind):
    import gpu_cases as G
    for name in list(G.CONV_CASES) + ["convT", "multi"]:
        try:
            run_one(kind, name)
        except Exception as e:  # noqa: BLE001
            print(json.dumps(dict(kind=kind, case=name, error=repr(e)[:300])))


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what == "one":
        return run_one(sys.argv[2], sys.argv[3])
    if what == "many":
        return run_many(sys.argv[2])
    import gpu_cases as G
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    lines = []
    kinds = ["simt", "tc"] if what == "all" else [what]
    if "simt" in kinds:
        kinds.remove("simt")
        r = subprocess.run([sys.executable, __file__, "many", "simt"], capture_output=True, text=True, timeout=600)
        for ln in r.stdout.strip().splitlines():
            print(ln, flush=True)
            lines.append(ln)
        if r.returncode != 0:
            lines.append("simt many rc=%d %s" % (r.returncode, r.stderr[-500:]))
            print(lines[-1], flush=True)
    for kind in kinds:
        for name in (list(G.HALO_CASES) if kind == "halo" else list(G.CONV_CASES) + ["convT", "multi"]):
            try:
                r = subprocess.run([sys.executable, __file__, "one", kind, name], capture_output=True, text=True,
                                   timeout=120)
                out = r.stdout.strip().splitlines()
                msg = out[-1] if out else ""
                if r.returncode != 0:
                    msg = "FAIL rc=%d %s | %s" % (r.returncode, msg, r.stderr.strip().splitlines()[-1:] )
            except subprocess.TimeoutExpired:
                msg = "TIMEOUT"
            line = "%s %s: %s" % (kind, name, msg)
            print(line, flush=True)
            lines.append(line)
    with open(os.path.join(ROOT, "gpurun_out", "diag.txt"), "w") as f:
        f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
