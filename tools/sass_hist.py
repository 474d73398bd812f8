#!/usr/bin/env python
"""Histogram of SASS opcodes per kernel of a .so/.cubin (cuobjdump -sass).  Usage: sass_hist.py lib.so [substr]"""
import re, subprocess, sys, collections
so = sys.argv[1]; pat = sys.argv[2] if len(sys.argv) > 2 else ""
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
cur = None; funcs = collections.OrderedDict()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m: cur = m.group(1); funcs[cur] = []; continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(@!?U?P\d\s+)?([A-Z0-9_]+)([.\w]*)", line)
    if m and cur: funcs[cur].append(m.group(2) + m.group(3))
for f, ops in funcs.items():
    if pat not in f: continue
    c = collections.Counter(o.split(".")[0] for o in ops)
    print(f, len(ops)); print("  " + "  ".join(f"{k}:{v}" for k, v in c.most_common(24)))
