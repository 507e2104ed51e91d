"""profiles/r02_sass_excerpt.txt: per kernel of the shipped .so, the counts of the SASS mnemonics that prove the tcgen05 / TMEM / TMA
design (B200_PROFILING.md): UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld), UTMALDG (TMA tensor load), UTCBAR (tcgen05.commit),
SYNCS (mbarrier), plus the first occurrence of each.  Usage: python scripts/make_sass_excerpt.py"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "revers_o_b200", "lib", "librevers_o_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
pat = re.compile(r"\b(UTCHMMA[.\w]*|LDTM[.\w]*|UTMALDG[.\w]*|UTCBAR[.\w]*|UTCATOMSWS[.\w]*|UTMAPF[.\w]*|SYNCS[.\w]*|ATOMG[.\w]*|ATOM[.\w]*|RED[.\w]*|STG[.\w]*|LDG[.\w]*)")
cur, per, first = None, collections.OrderedDict(), {}
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        per[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    for tok in pat.findall(line):
        base = tok if tok.startswith(("UTC", "LDTM", "UTMA")) else tok.split(".")[0]
        per[cur][base] += 1
        if tok.startswith(("UTC", "LDTM", "UTMA")) and (cur, base) not in first:
            first[(cur, base)] = line.strip()
out = ["# SASS evidence of the shipped library (cuobjdump -sass revers_o_b200/lib/librevers_o_b200.so, sm_100a)", "",
       "mnemonic counts per kernel; tensor-core / TMEM / TMA mnemonics listed with their first occurrence", ""]
for k, c in per.items():
    tc = {m: n for m, n in c.items() if m.startswith(("UTC", "LDTM", "UTMA"))}
    other = {m: n for m, n in c.items() if not m.startswith(("UTC", "LDTM", "UTMA"))}
    out.append(f"## {k}")
    out.append("  tcgen05/TMEM/TMA: " + (", ".join(f"{m} x{n}" for m, n in sorted(tc.items())) or "none (CUDA-core kernel)"))
    out.append("  memory/sync:      " + ", ".join(f"{m} x{n}" for m, n in sorted(other.items())))
    for m in sorted(tc):
        out.append("      " + first[(k, m)])
    out.append("")
open(os.path.join(ROOT, "profiles", "r02_sass_excerpt.txt"), "w").write("\n".join(out))
print("kernels:", len(per))
