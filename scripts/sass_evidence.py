#!/usr/bin/env python
"""profiles/r02_sass_tcgen05_evidence.txt: per kernel of the built library, the Blackwell-native SASS mnemonics
(UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG = cp.async.bulk.tensor, UTCBAR = tcgen05.commit,
SYNCS = mbarrier) counted from `cuobjdump -sass` -- and HMMA / HGMMA (legacy tensor paths), which must stay at zero."""
import collections
import re
import subprocess
import sys

sys.path.insert(0, ".")
from immunostruct_b200.build import LIB_PATH  # noqa: E402

sass = subprocess.run(["cuobjdump", "-sass", LIB_PATH], capture_output=True, text=True).stdout
names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
MN = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "HGMMA"]
rows, cur, counts = [], None, None
it = iter(names)
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        if cur:
            rows.append((cur, counts))
        cur, counts = next(it), collections.Counter()
        continue
    if cur:
        for k in MN:
            if re.search(rf"\b{k}\b|\b{k}\.", line):
                counts[k] += 1
if cur:
    rows.append((cur, counts))
out = ["# SASS evidence (cuobjdump -sass immunostruct_b200/libimmunostruct_b200.so), one row per kernel that uses the",
       "# tensor cores, TMEM, TMA or mbarriers.  UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG =",
       "# cp.async.bulk.tensor (TMA load), UTCBAR = tcgen05.commit, SYNCS = mbarrier ops.  HMMA / HGMMA (mma.sync / wgmma) = 0.",
       "", "| kernel | " + " | ".join(MN) + " |", "|---|" + "---:|" * len(MN)]
tot = collections.Counter()
for name, c in sorted(rows):
    tot.update(c)
    if any(c[k] for k in MN):
        short = (re.sub(r">\(.*$", ">", name) if ">(" in name else re.sub(r"\(.*$", "", name))[:110]
        out.append(f"| `{short}` | " + " | ".join(str(c[k]) for k in MN) + " |")
out += ["", f"kernels in the library: {len(rows)}; totals: " + ", ".join(f"{k} {tot[k]}" for k in MN)]
open("profiles/r02_sass_tcgen05_evidence.txt", "w").write("\n".join(out) + "\n")
print("\n".join(out[-3:]))
print(len([r for r in rows if r[1]["UTCHMMA"]]), "kernels with UTCHMMA;", len([r for r in rows if r[1]["UTMALDG"]]), "with UTMALDG")
