#!/usr/bin/env python
"""Opcode histogram of libpcfm_b200.so (cuobjdump -sass), per kernel family and in total:
    python profiles/sass_opcodes.py > profiles/r02_sass_opcodes.txt
Evidence for the Blackwell-specific instructions the kernels rely on: UBLKCP (1-D TMA bulk copies), FFMA2 / FADD2 /
FMUL2 (packed fp32), UCGABAR_* (cluster barriers), VIMNMX3, REDUX / CREDUX, ATOMS / REDG, MUFU.EX2."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "point-cloud-flow-matching_b200", "lib", "libpcfm_b200.so")
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
fam, per, total = None, collections.defaultdict(collections.Counter), collections.Counter()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        fam = re.sub(r"<.*", "", name.split("(")[0]).replace("void ", "").replace("pcfm::", "")
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)*)", line)
    if m and fam:
        op = m.group(1)
        base = op.split(".")[0]
        key = op if base in ("MUFU", "ATOMS", "REDG", "ATOMG", "RED", "UBLKCP", "REDUX", "CREDUX", "SYNCS") else base
        key = ".".join(key.split(".")[:3])
        per[fam][key] += 1
        total[key] += 1
WATCH = ["UBLKCP", "FFMA2", "FADD2", "FMUL2", "UCGABAR_ARV", "UCGABAR_WAIT", "VIMNMX3", "REDUX", "CREDUX", "MUFU.EX2", "MUFU.RCP",
         "ATOMS", "REDG", "ATOMG", "ACQBULK", "UTMALDG", "UTCHMMA", "HMMA"]
print(f"# {os.path.relpath(LIB, ROOT)}: SASS opcode counts (static instructions), cuobjdump -sass")
print("# total:", sum(total.values()), "instructions in", len(per), "kernel families")
def pick(cnt):
    res = []
    for w in WATCH:
        n = sum(v for k, v in cnt.items() if k == w or k.startswith(w + "."))
        if n:
            res.append(f"{w} {n}")
    return ", ".join(res)
print("ALL KERNELS:", pick(total))
print("absent (as expected: no tensor-map TMA, no tensor-core contraction on this path):",
      ", ".join(w for w in ("UTMALDG", "UTCHMMA", "HMMA") if not any(k.startswith(w) for k in total)))
for f in sorted(per):
    print(f"{f:42s} {sum(per[f].values()):6d} instr | {pick(per[f])}")
