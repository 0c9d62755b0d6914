"""ncu launch list -> per-kernel share table of ONE forward (the slice between two image conversions).
usage: python tools/launch_summary.py launches.csv out.md "<command line>" """
import collections
import csv
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
L = [(r[ik].split("(")[0].replace("void <unnamed>::", "").replace("<unnamed>::", "").replace("void ", "").strip(),
      float(r[iv])) for r in rows[1:]]
starts = [i for i, (k, _) in enumerate(L) if k.startswith("planar_to_cl_kernel")]
assert len(starts) >= 2, "capture window does not contain a whole forward"
seq = L[starts[-2]:starts[-1]]
agg = collections.defaultdict(lambda: [0, 0.0])
for k, t in seq:
    agg[k][0] += 1
    agg[k][1] += t
tot = sum(v[1] for v in agg.values())
with open(sys.argv[2], "w") as f:
    f.write("# ncu launch list of one forward (config 2, CUDA graph off)\n\n")
    f.write("command: `%s`\n\n" % sys.argv[3])
    f.write("Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.  "
            "%d launches, %.3f ms summed.\n\n" % (len(seq), tot / 1e6))
    f.write("| kernel | launches | sum ms | share |\n|---|---|---|---|\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write("| `%s` | %d | %.3f | %.1f%% |\n" % (k, v[0], v[1] / 1e6, 100 * v[1] / tot))
print(open(sys.argv[2]).read())
