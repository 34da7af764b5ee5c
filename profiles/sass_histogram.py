"""Per-kernel SASS opcode histogram of libstego_b200.so (run anywhere with cuobjdump, no GPU needed):
    python profiles/sass_histogram.py > profiles/r2_sass_histogram.md
Shows which kernels carry tcgen05 (UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM / STTM = tcgen05.ld / st) and TMA
(UTMALDG / UTMASTG / UTMAREDG / UTMAPF) instructions, and that no legacy HMMA (mma.sync) is anywhere."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "stego_b200", "libstego_b200.so")
KEY = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTMAPF", "SYNCS", "HMMA", "MUFU", "FFMA2", "RED", "ATOM",
       "STL", "LDL"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
kern, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern).replace("stego::", "")
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and kern:
        op = m.group(2)
        hist[kern][op] += 1
        hist[kern]["_total"] += 1
print(f"SASS opcode counts per kernel of `stego_b200/libstego_b200.so` (`cuobjdump -sass`, static instruction counts; sm_100a)\n")
print("| kernel | instr | " + " | ".join(KEY) + " |")
print("|---|---|" + "---|" * len(KEY))
tot = collections.Counter()
for k, c in hist.items():
    cells = []
    for key in KEY:
        n = sum(v for op, v in c.items() if op.startswith(key))
        tot[key] += n
        cells.append(str(n) if n else "")
    print(f"| `{k[:70]}` | {c['_total']} | " + " | ".join(cells) + " |")
print("| **all kernels** | | " + " | ".join(str(tot[k]) for k in KEY) + " |")
