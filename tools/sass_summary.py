"""SASS evidence for profiles/: per kernel of libslsb200.so, the instruction count and how many are DMMA (FP64 tensor),
DFMA, SHFL, LDS/STS, LDG/STG, BAR — and the first DMMA lines of each solve kernel.
    python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "systemlevelcontrol.jl_b200", "lib", "libslsb200.so")
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}  (sm_100a)")
kern = None
ops = collections.OrderedDict()
first = {}
for ln in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        ops[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m and kern:
        op = m.group(1)
        ops[kern][op.split(".")[0]] += 1
        if op.startswith("DMMA") and len(first.setdefault(kern, [])) < 3:
            first[kern].append(re.sub(r"/\* 0x[0-9a-f]+ \*/", "", ln).rstrip())
cols = ["DMMA", "DFMA", "DMUL", "DADD", "MUFU", "SHFL", "LDS", "STS", "LDG", "STG", "BAR", "ATOMS", "ATOMG"]
print(f"{'kernel':70s} {'total':>7s} " + " ".join(f"{c:>6s}" for c in cols))
for k, c in ops.items():
    name = k if len(k) <= 70 else k[:67] + "..."
    print(f"{name:70s} {sum(c.values()):7d} " + " ".join(f"{c.get(x, 0):6d}" for x in cols))
print("\n# first DMMA instructions of the solve kernels (FP64 tensor-core MMA m8n8k4)")
for k, lines in first.items():
    print(k)
    for l in lines:
        print("   ", l.strip())
