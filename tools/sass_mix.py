#!/usr/bin/env python
"""Static instruction mix of the main loop of a kernel from `cuobjdump -sass` output.

usage: cuobjdump -sass -fun <mangled> obj.o | python tools/sass_mix.py
The main loop is taken as the longest backward branch.  Prints instructions per loop trip by pipe
class, so that a change to a kernel can be judged here (no GPU) before it is timed on one."""
import collections
import re
import sys

ins = []
for line in sys.stdin:
    m = re.match(r"\s*/\*([0-9a-f]{4,6})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
best = None
for a, t in ins:
    m = re.search(r"\bBRA(?:\.\w+)*\s+(?:\S+,\s*)?(0x[0-9a-f]+)", t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt < a and (best is None or a - tgt > best[1] - best[0]):
            best = (tgt, a)
if best is None:
    sys.exit("no backward branch found")
body = [t for a, t in ins if best[0] <= a <= best[1]]


def klass(t):
    op = re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0]
    if op in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"):
        return "fp64"
    if op in ("MUFU", "F2F", "I2F", "F2I", "I2FP", "F2FP"):
        return "xu"
    if op in ("LDS", "STS", "LDG", "STG", "LDGSTS", "LDSM", "ATOMS", "ATOMG", "RED", "LDL", "STL", "LDC", "LDCU", "SHFL"):
        return "lsu/" + op
    if op in ("BAR", "BRA", "BSSY", "BSYNC", "EXIT", "WARPSYNC", "DEPBAR", "LDGDEPBAR", "NOP", "CALL", "RET"):
        return "ctl/" + op
    if op in ("IMAD", "FFMA", "FMUL", "FADD", "HFMA2"):
        return "fma/" + op
    return "alu/" + op


c = collections.Counter(klass(t) for t in body)
tot = len(body)
print("loop 0x%x..0x%x: %d instructions" % (best[0], best[1], tot))
groups = collections.Counter()
for k, v in c.items():
    groups[k.split("/")[0]] += v
for g, v in groups.most_common():
    print("  %-5s %5d  %4.1f %%" % (g, v, 100.0 * v / tot))
    for k, n in sorted(((k, n) for k, n in c.items() if k.startswith(g + "/")), key=lambda x: -x[1]):
        print("      %-12s %5d" % (k.split("/")[1], n))
