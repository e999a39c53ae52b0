"""Per-kernel counts of the SASS mnemonics that prove a Blackwell-native kernel (B200_PROFILING.md): UTC*MMA = tcgen05.mma,
LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG / UTMAREDG / UBLKCP = TMA, HMMA = the legacy mma.sync path (must be absent).
  python tools/sass_summary.py > profiles/r2_sass_summary.txt        (cuobjdump on the in-tree library; no GPU needed)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "vision_pt_b200", "libvptb200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.splitlines()
KEYS = ("UTCHMMA", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "LDTM", "STTM", "HMMA", "MUFU")
rows, cur, it = [], None, iter(names)
counts = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        if cur is not None:
            rows.append((cur, counts))
        cur, counts = next(it), collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", line)
    if m and counts is not None:
        op = m.group(1)
        for k in KEYS:
            if op.startswith(k):
                counts[k] += 1
        counts["_total"] += 1
if cur is not None:
    rows.append((cur, counts))
print(f"# cuobjdump -sass vision_pt_b200/libvptb200.so ({os.path.getsize(lib) / 1e6:.1f} MB, sm_100a): mnemonic counts per kernel")
print(f"# {'kernel':86s} " + " ".join(f"{k:>8s}" for k in KEYS) + "   instrs")
tot = collections.Counter()
for name, c in sorted(rows, key=lambda r: -r[1]["UTCHMMA"]):
    short = re.sub(r"\(.*", "", name).replace("vpt::", "").replace("void ", "")[:86]
    print(f"{short:88s} " + " ".join(f"{c[k]:8d}" for k in KEYS) + f" {c['_total']:8d}")
    tot.update(c)
print(f"{'TOTAL (' + str(len(rows)) + ' kernels)':88s} " + " ".join(f"{tot[k]:8d}" for k in KEYS) + f" {tot['_total']:8d}")
assert tot["HMMA"] == 0, "legacy mma.sync found"
