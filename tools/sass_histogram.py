#!/usr/bin/env python3
"""SASS opcode histogram of one kernel of a built object (cuobjdump -sass), written as a markdown table.
  python tools/sass_histogram.py <object> <kernel substring> <out.md>"""
import collections
import re
import subprocess
import sys

obj, kern, out = sys.argv[1], sys.argv[2], sys.argv[3]
text = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
hist = collections.Counter()
cur = None
arch = re.search(r"arch = (\S+)", text)
for line in text.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    if cur and kern in cur:
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m:
            hist[m.group(1)] += 1
total = sum(hist.values())
lines = [f"# SASS opcode histogram: `{kern}` in `{obj}` ({arch.group(1) if arch else '?'})", "",
         f"{total} instructions (static count; `cuobjdump -sass`).  The field multiplier is `IMAD.WIDE.U32` with and without the carry flag (`.X`): "
         "one 32x32->64 multiply-accumulate each, the LMAC of the roofline.  No tensor-core or TMA opcodes: nothing on this path is a contraction.", "",
         "| opcode | count | share |", "|---|---:|---:|"]
for op, c in hist.most_common(40):
    lines.append(f"| `{op}` | {c} | {100 * c / total:.1f}% |")
imad = sum(c for op, c in hist.items() if op.startswith("IMAD.WIDE.U32"))
lines += ["", f"`IMAD.WIDE.U32*` together: {imad} ({100 * imad / total:.1f}% of the static instructions)."]
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:24]))
