"""Per-kernel SASS mnemonic counts of libttlda.so (cuobjdump -sass): which kernels use the FP64 tensor core (DMMA), the
TMA engine (UTMALDG), global reductions (REDG), 16-byte global / generic / shared loads.

    python tools/sass_summary.py [tuotuo_b200/libttlda.so] > profiles/r02_sass_summary.txt
"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "tuotuo_b200/libttlda.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
PATS = {"DMMA": r"\bDMMA", "UTMALDG": r"\bUTMALDG", "REDG.F64": r"\bREDG\.E\.ADD\.F64", "LDG.128": r"\bLDG\.E\.128",
        "LD.128(generic)": r"\bLD\.E\.128", "LDS.128": r"\bLDS\.128", "STG.128": r"\bSTG\.E\.128", "MUFU.RCP64H": r"MUFU\.RCP64H",
        "DFMA": r"\bDFMA", "SHFL": r"\bSHFL", "ATOMG": r"\bATOMG", "BAR": r"\bBAR\.SYNC", "LDGSTS": r"\bLDGSTS", "ATOMS": r"\bATOMS"}
cur, counts, arch = None, collections.OrderedDict(), set()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        counts.setdefault(cur, collections.Counter())
        continue
    m = re.search(r"arch = (sm_\w+)", line)
    if m:
        arch.add(m.group(1))
    if cur:
        for k, p in PATS.items():
            if re.search(p, line):
                counts[cur][k] += 1
print(f"# {lib}: cubins for {sorted(arch)}; {len(counts)} kernels")
print(f"# {'kernel':78s} " + " ".join(f"{k:>8s}" for k in PATS))
tot = collections.Counter()
for name, c in counts.items():
    if not any(c.values()):
        continue
    tot.update(c)
    print(f"{name[:80]:80s} " + " ".join(f"{c[k]:8d}" for k in PATS))
print(f"{'TOTAL':80s} " + " ".join(f"{tot[k]:8d}" for k in PATS))
