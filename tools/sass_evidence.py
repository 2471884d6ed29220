"""cuobjdump -sass of the built library -> profiles/r02_sass_evidence.md: per kernel, the counts of the tcgen05 / TMA / mbarrier / mma.sync
mnemonics (UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor, SYNCS = mbarrier, HMMA = mma.sync,
ELECT = elect.sync, BRA.U.ANY = the per-instruction loop nvcc emits around uniform-datapath instructions in a divergent branch: 0 now)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "gan_deeplearning4j_b200", "lib", "libb200gan.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
MN = ["UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "SYNCS", "HMMA", "LDSM", "ELECT", "BRA.U.ANY", "UTCATOMSWS", "ACQBULK", "CCTL"]
cur, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        cur = cur.replace("void b2g::", "").replace("b2g::", "")
        counts.setdefault(cur, collections.Counter())
        continue
    if cur:
        for k in MN:
            if re.search(r"\b" + re.escape(k) + r"\b", line):
                counts[cur][k] += 1
rows = [(k, c) for k, c in counts.items() if c["UTCHMMA"] or c["UTMALDG"] or c["HMMA"]]
md = ["# SASS evidence (round 2): `cuobjdump -sass gan_deeplearning4j_b200/lib/libb200gan.so`, instruction counts per kernel instantiation", "",
      "UTCHMMA = `tcgen05.mma`, UTCBAR = `tcgen05.commit`, LDTM = `tcgen05.ld`, UTMALDG = `cp.async.bulk.tensor` (TMA), SYNCS = mbarrier ops, HMMA / LDSM = `mma.sync` / `ldmatrix`",
      "(the two G-first kernels), ELECT = `elect.sync`.  BRA.U.ANY = the loop nvcc wraps around a uniform-datapath instruction inside a divergent",
      "single-lane branch: zero everywhere since the roles became converged warps + elect.sync.", "",
      "| kernel | " + " | ".join(MN[:9]) + " |", "|---|" + "---|" * 9]
for k, c in rows:
    md.append(f"| `{k[:110]}` | " + " | ".join(str(c[m]) for m in MN[:9]) + " |")
open(os.path.join(ROOT, "profiles", "r02_sass_evidence.md"), "w").write("\n".join(md) + "\n")
print(len(rows), "kernels;", sum(c["BRA.U.ANY"] for _, c in rows), "BRA.U.ANY in tensor kernels")
