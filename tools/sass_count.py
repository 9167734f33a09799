#!/usr/bin/env python3
"""Counts SASS instructions per kernel of a cubin/.so (instruction-cache footprint check)."""
import re, subprocess, sys
out = subprocess.check_output(["cuobjdump", "-sass", sys.argv[1]], text=True)
name, n = None, {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1); n[name] = 0; continue
    if name and re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", line):
        n[name] += 1
for k, v in sorted(n.items(), key=lambda kv: -kv[1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 10]:
    print(f"{v:7d} instr  {v*16/1024:7.1f} KB  {k[:110]}")
