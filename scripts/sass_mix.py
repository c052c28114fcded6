"""Static SASS instruction mix of one kernel of libpercnn_b200.so (cuobjdump), e.g.
    python scripts/sass_mix.py bwd_tmaILi0ELb0ELb1      # substring of the mangled name
Used on the CPU box to compare instruction counts of kernel variants before spending GPU time."""
import collections
import re
import subprocess
import sys

pat = sys.argv[1]
lib = sys.argv[2] if len(sys.argv) > 2 else "percnn_b200/libpercnn_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout.split("\n")
start = next(i for i, l in enumerate(txt) if "Function :" in l and pat in l)
end = next((i for i in range(start + 1, len(txt)) if "Function :" in txt[i]), len(txt))
ops = collections.Counter()
n = 0
for l in txt[start:end]:
    m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
    if m:
        ops[m.group(2).split(".")[0]] += 1
        n += 1
print(txt[start].strip(), "->", n, "instructions")
print("  ".join(f"{k}:{v}" for k, v in ops.most_common(28)))
# hot body estimate: instructions from the first LDS.128 to the last STG.E.128 of the function
body = []
inside = False
idx = [i for i, l in enumerate(txt[start:end]) if "STG.E.128" in l]
first = next((i for i, l in enumerate(txt[start:end]) if "LDS.128" in l), None)
if not idx or first is None:      # kernels without 128-bit loads/stores: the whole-function mix above is all there is
    sys.exit(0)
bops = collections.Counter()
nb = 0
for l in txt[start + first:start + idx[-1] + 1]:
    m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
    if m:
        bops[m.group(2).split(".")[0]] += 1
        nb += 1
print(f"first LDS.128 .. last STG.E.128: {nb} instructions, {len(idx)} STG.E.128")
print("  ".join(f"{k}:{v}" for k, v in bops.most_common(24)))
