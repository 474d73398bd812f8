#!/usr/bin/env python
"""Dump the SASS of one kernel (substring match) from a .so: address, opcode+operands."""
import re, subprocess, sys
so, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
cur = None
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); continue
    if cur and pat in cur:
        m = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+(.*?)\s*/\*", line)
        if m: print(m.group(1), m.group(2))
