"""Per-kernel counts of the SASS mnemonics that prove (or disprove) tcgen05 / TMEM / TMA use -- from the built library, no GPU needed.

    python scripts/sass_summary.py > profiles/r2_sass_summary.txt

UTCHMMA = tcgen05.mma (kind::f16)   LDTM = tcgen05.ld   UTCBAR = tcgen05.commit   STSM = stmatrix
UBLKCP = cp.async.bulk (the TMA engine's 1-D bulk copy)   UBLKPF = cp.async.bulk.prefetch.L2
UTMALDG / UTMASTG = cp.async.bulk.tensor (tiled TMA) load / store   HMMA = legacy mma.sync   FFMA / DFMA = CUDA-core fp32 / fp64 FMA
"""
import collections
import os
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "heybuddy_b200", "_lib", "libheybuddy_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
keys = ["UTCHMMA", "LDTM", "UTCBAR", "STSM", "UBLKCP", "UBLKPF", "UTMALDG", "UTMASTG", "HMMA", "FFMA", "DFMA"]
counts, name = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        counts[name] = collections.Counter()
        continue
    if name is None:
        continue
    for k in keys:
        if re.search(r"\b" + k + r"\b|\b" + k + r"\.", line) and not (k == "HMMA" and "UTCHMMA" in line):
            counts[name][k] += 1
names = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
print(__doc__.strip().split("\n\n", 1)[1])
print()
print(f"{'kernel':<78}" + "".join(f"{k:>8}" for k in keys))
rows = []
for mangled, pretty in zip(counts, names):
    pretty = re.sub(r"\(anonymous namespace\)::", "", pretty.replace("hb::", "").replace("void ", ""))
    pretty = re.sub(r"\(.*", "", pretty)
    rows.append((pretty[:76], counts[mangled]))
for pretty, c in sorted(rows):
    print(f"{pretty:<78}" + "".join(f"{c[k]:>8}" for k in keys))
tot = collections.Counter()
for _, c in rows:
    tot.update(c)
print(f"{'TOTAL':<78}" + "".join(f"{tot[k]:>8}" for k in keys))
