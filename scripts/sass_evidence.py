"""Per-kernel SASS evidence from the built library: instruction counts that show which hardware paths a kernel uses
(UTMALDG/UTMASTG/UBLKCP = TMA, SYNCS = mbarrier, FFMA2/FMUL2 = packed fp32x2, LDS.128/STG.128/LDG.128 = 128-bit accesses,
USETMAXREG = register reallocation, BAR.ARV = non-blocking named-barrier arrive, RED/ATOM ... .SYS = cross-GPU flags).

    python scripts/sass_evidence.py > profiles/r02_sass_evidence.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "percnn_b200", "libpercnn_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pats = [("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"), ("UBLKCP", r"\bUBLKCP"), ("SYNCS", r"\bSYNCS"), ("FFMA2", r"\bFFMA2"),
        ("FMUL2", r"\bFMUL2"), ("FFMA", r"\bFFMA\b"), ("DFMA", r"\bDFMA"), ("LDS.128", r"\bLDS\.128"), ("STS.128", r"\bSTS\.128"),
        ("LDG.128", r"\bLDG\.E\.128"), ("STG.128", r"\bSTG\.E\.128"), ("SHFL", r"\bSHFL"), ("USETMAXREG", r"USETMAXREG"),
        ("BAR.ARV", r"BAR\.ARV"), ("sys-scope", r"\.SYS\b"), ("ACQBULK/PREEXIT", r"ACQBULK|PREEXIT"), ("STL/LDL", r"\b(STL|LDL)\b")]
cur, counts, total = None, collections.OrderedDict(), {}
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        cur = cur.replace("percnn::", "").replace("void ", "")
        counts[cur] = collections.Counter()
        total[cur] = 0
        continue
    if cur and re.match(r"\s+/\*[0-9a-f]{4,6}\*/", line):
        total[cur] += 1
        for name, pat in pats:
            if re.search(pat, line):
                counts[cur][name] += 1
print(f"{'kernel':58s} {'instr':>6s} " + " ".join(f"{n:>8s}" for n, _ in pats))
for k in sorted(counts):
    if re.search(r"<[1-5]\b|<[1-5],", k):        # slots 1..5 are copies of slot 0
        continue
    print(f"{k[:58]:58s} {total[k]:6d} " + " ".join(f"{counts[k][n]:8d}" for n, _ in pats))
