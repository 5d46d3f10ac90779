#!/usr/bin/env python
"""Per-kernel counts of the Blackwell-specific SASS opcodes in libuahn.so (the evidence table of B200_PROFILING.md):

    python tools/sass_opcodes.py > profiles/r02_sass_opcodes.txt

tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UTMASTG/UBLKCP, tcgen05.commit -> UTCBAR,
TMEM alloc -> UTCATOMSWS / UTCALLOC-style ops, cp.async -> LDGSTS; texture gather -> TLD4, surface store -> SUST; the legacy
tensor path (mma.sync -> HMMA, ldmatrix -> LDSM) is used by the batch-1 split-K kernel only.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "cuahn_vio_b200", "lib", "libuahn.so")
WATCH = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "UTCATOMSWS", "UTCCP", "LDGSTS",
         "HMMA", "UCGABAR", "TLD4", "SUST", "LDSM"]

sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
counts, order, cur, total = {}, [], None, collections.Counter()
it = iter(names)
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = next(it)
        cur = re.sub(r"\(anonymous namespace\)::|uahn::", "", cur)
        cur = cur.split("(")[0] if "<" not in cur else cur[:cur.rindex(">") + 1] if ">(" in cur else cur
        counts[cur] = collections.Counter()
        order.append(cur)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", line)
    if m and cur:
        op, mods = m.group(1), m.group(2)
        counts[cur]["_all"] += 1
        if op in WATCH:
            counts[cur][op] += 1
            total[op] += 1
            if op in ("UTCHMMA", "UTMALDG", "UTCBAR") and mods:
                counts[cur][op + mods] += 1
print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)} — Blackwell opcode counts per kernel (static instruction counts)")
print("# totals: " + ", ".join(f"{k} {v}" for k, v in sorted(total.items())))
for k in order:
    c = counts[k]
    hot = {o: n for o, n in c.items() if o != "_all"}
    if not hot:
        continue
    print(f"{k}\n    instructions {c['_all']}: " + ", ".join(f"{o} {n}" for o, n in sorted(hot.items())))
