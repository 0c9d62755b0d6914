"""Join an ncu launch list (gpu__time_duration) of one forward with the plan's op names.
usage: python tools/ncu_by_op.py gpurun_out/launches.csv gpurun_out/plan_profile.json"""
import collections
import csv
import json
import re
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
launches = [(r[ik].split("(")[0].replace("void <unnamed>::", "").strip(), float(r[iv]) / 1e3) for r in rows[1:]]
ops = json.load(open(sys.argv[2]))


def nk(name):
    return 2 if name.endswith(".se") else 1


need = sum(nk(o["name"]) for o in ops)
# find the LAST complete forward: it starts with planar_to_cl (image) and the ops' kernels follow
starts = [i for i, (k, _) in enumerate(launches) if "planar_to_cl_kernel" in k]
full = [s for s in starts if s + 1 + need <= len(launches)]
if full:
    seq = launches[full[-1] + 1: full[-1] + 1 + need]
else:
    # align from the END of the forward that precedes the first planar_to_cl (trailing cl_to_planar block)
    end = starts[0]
    while end > 0 and "cl_to_planar" in launches[end - 1][0]:
        end -= 1
    miss = max(0, need - end)
    seq = [("missing", 0.0)] * miss + launches[max(0, end - need): end]
    print("note: first %d kernels of the forward were outside the capture window" % miss)
cat = collections.defaultdict(lambda: [0, 0.0])
per = []
i = 0
for o in ops:
    n = o["name"]
    t = sum(seq[i + j][1] for j in range(nk(n)))
    i += nk(n)
    per.append((n, t))
    if n.startswith("head."): k = "3d.head"
    elif re.match(r"b\d+\.\d+\.dw", n): k = "enc.dw"
    elif re.match(r"b\d+\.\d+\.se", n): k = "enc.se+fold"
    elif re.match(r"b\d+\.\d+\.(expand|proj)", n) or n in ("stem", "conv_head"): k = "enc.1x1"
    elif "bilinear" in n: k = "dec.bilinear"
    elif re.match(r"up\d+\.conv", n): k = "dec.conv"
    elif n.startswith("resize_") or n == "dec.conv2": k = "dec.1x1"
    elif n.startswith("sfa"): k = "lift"
    else: k = "3d.other"
    cat[k][0] += 1
    cat[k][1] += t
tot = sum(v[1] for v in cat.values())
print("one forward under ncu (cold-cache, serialised): %.3f ms over %d kernels" % (tot / 1e3, need))
for k, v in sorted(cat.items(), key=lambda kv: -kv[1][1]):
    print("%-14s n=%4d %8.3f ms %5.1f%%" % (k, v[0], v[1] / 1e3, 100 * v[1] / tot))
st = collections.defaultdict(lambda: collections.defaultdict(float))
for n, t in per:
    m = re.match(r"b(\d+)\.(\d+)\.(\w+)", n)
    if m:
        st[int(m.group(1))][m.group(3)] += t
print("stage   expand      dw      se    proj  (us)")
for s_ in sorted(st):
    d = st[s_]
    print("b%d   %8.1f %8.1f %8.1f %8.1f" % (s_, d.get("expand", 0), d["dw"], d["se"], d["proj"]))
per.sort(key=lambda x: -x[1])
for n, t in per[:25]:
    print("   %-28s %8.1f us" % (n, t))
