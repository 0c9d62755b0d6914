"""SASS evidence: per-kernel counts of the tcgen05 / TMA / mbarrier instructions in the built library.
usage: python tools/sass_listing.py > profiles/<round>_sass_conv.txt   (cuobjdump on PATH, library built)"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "occdepth_b200", "lib", "libocc_b200.so")
MNEMONICS = ("UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "LDTM", "UTCBAR", "SYNCS", "LDGSTS", "UBLKCP")
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
filt = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", out)), capture_output=True,
                      text=True).stdout.splitlines()
names = iter(filt)
counts, cur = collections.OrderedDict(), None
for line in out.splitlines():
    if "Function : " in line:
        cur = next(names)
        cur = cur.replace("void ", "").replace("(anonymous namespace)::", "")
        cur = re.sub(r"\(.*", "", cur)
        counts[cur] = collections.Counter()
    elif cur:
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?\w+\s+)?([A-Z0-9_]+)", line)
        if m and m.group(1) in MNEMONICS:
            counts[cur][m.group(1)] += 1
print("# SASS evidence (cuobjdump -sass occdepth_b200/lib/libocc_b200.so, sm_100a): instruction counts per kernel.")
print("# UTCHMMA = tcgen05.mma (kind::f16 and kind::tf32 both disassemble to UTCHMMA; the operand type sits in the")
print("# instruction descriptor register), UTMALDG = TMA tensor load (cp.async.bulk.tensor), LDTM = tcgen05.ld,")
print("# UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, LDGSTS = cp.async (depthwise tile staging).")
print("# Template arguments of the conv kernels: <element type (float = TF32 operands), K-chunk row bytes[, x-packed]>")
for k, c in counts.items():
    if c:
        print("%-70s %s" % (k, "  ".join("%s=%d" % kv for kv in sorted(c.items()))))
