#!/usr/bin/env python
"""Compact table of `nvcc -Xptxas -v` output: kernel, registers, spill bytes, stack, smem."""
import re
import subprocess
import sys

log = open(sys.argv[1]).read() if len(sys.argv) > 1 else sys.stdin.read()
pat = re.compile(r"Compiling entry function '(\S+)'.*?\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, "
                 r"(\d+) bytes spill loads\n.*?Used (\d+) registers", re.S)
rows = []
for m in pat.finditer(log):
    name = m.group(1)
    try:
        name = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    except Exception:
        pass
    name = re.sub(r"\(.*", "", name).replace("void msed::", "")
    rows.append((name, int(m.group(5)), int(m.group(3)), int(m.group(4)), int(m.group(2))))
only = sys.argv[2] if len(sys.argv) > 2 else ""
for r in rows:
    if only in r[0]:
        print(f"{r[0]:60s} regs={r[1]:3d} spill_st={r[2]:4d} spill_ld={r[3]:4d} stack={r[4]}")
