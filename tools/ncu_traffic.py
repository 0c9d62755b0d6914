"""ncu CSV (gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum per launch of ONE forward) ->
profiles/ncu_traffic.json (what bench.py quotes as roofline.traffic) + a per-kernel markdown table.
usage: python tools/ncu_traffic.py <precision> <ncu.csv> <out.md>"""
import collections
import csv
import json
import os
import re
import sys

prec, path, out_md = sys.argv[1], sys.argv[2], sys.argv[3]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
hdr = rows[0]
ik, im, iv, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
iu = hdr.index("Metric Unit")
launch = collections.OrderedDict()
for r in rows[1:]:
    d = launch.setdefault(r[iid], {"name": re.sub(r"\(.*", "", r[ik]).replace("void ", "").replace("(anonymous namespace)::", "")
                                   .replace("<unnamed>::", "").strip()})
    v = float(r[iv].replace(",", ""))
    u = r[iu].lower()
    if "byte" in u:
        v *= {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    if "time" in r[im]:
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)      # -> ms
    d[r[im]] = v
agg = collections.OrderedDict()
for d in launch.values():
    a = agg.setdefault(d["name"], [0, 0.0, 0.0, 0.0])
    a[0] += 1
    a[1] += d.get("gpu__time_duration.sum", 0.0)
    a[2] += d.get("dram__bytes_read.sum", 0.0)
    a[3] += d.get("dram__bytes_write.sum", 0.0)
tot_ms = sum(a[1] for a in agg.values())
conv = [a for k, a in agg.items() if k.startswith("conv_tc_kernel") or k.startswith("conv_halo")]
lift = [a for k, a in agg.items() if k.startswith("sfa_lift")]
res = {"conv_family_dram_bytes": sum(a[2] + a[3] for a in conv), "conv_family_launches": sum(a[0] for a in conv),
       "conv_family_ms_under_ncu": sum(a[1] for a in conv),
       "lift_dram_bytes": sum(a[2] + a[3] for a in lift), "forward_dram_bytes": sum(a[2] + a[3] for a in agg.values()),
       "forward_ms_under_ncu": tot_ms, "launches": sum(a[0] for a in agg.values()),
       "source": os.path.basename(path)}
jp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
allj = json.load(open(jp)) if os.path.exists(jp) else {}
allj[prec] = res
json.dump(allj, open(jp, "w"), indent=1)
with open(out_md, "w") as f:
    f.write("# ncu launch list of one config-2 forward, precision mode %s (CUDA graph off)\n\n" % prec)
    f.write("`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none "
            "--profile-from-start off python tools/ncu_forward.py %s`\n\n" % prec)
    f.write("Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.  %d launches, "
            "%.3f ms summed, %.2f GB of DRAM traffic.\n\n" % (res["launches"], tot_ms, res["forward_dram_bytes"] / 1e9))
    f.write("| kernel | launches | sum ms | share | DRAM read MB | DRAM write MB |\n|---|---|---|---|---|---|\n")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write("| `%s` | %d | %.3f | %.1f%% | %.1f | %.1f |\n" % (k, a[0], a[1], 100 * a[1] / tot_ms, a[2] / 1e6, a[3] / 1e6))
print(open(out_md).read())
print(json.dumps(res, indent=1))
