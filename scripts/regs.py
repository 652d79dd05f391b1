#!/usr/bin/env python
"""Registers / spills per kernel from ptxas -v output (sextans_b200/csrc/ptxas.log)."""
import re, subprocess, sys
log = open(sys.argv[1] if len(sys.argv) > 1 else "sextans_b200/csrc/ptxas.log").read()
pat = sys.argv[2] if len(sys.argv) > 2 else ""
cur = None
for line in log.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        continue
    m = re.search(r"Used (\d+) registers", line)
    if m and cur and pat in cur:
        print(f"{int(m.group(1)):4d}  {cur}")
    m = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m and cur and pat in cur and (int(m.group(1)) or int(m.group(2))):
        print(f"      SPILL {m.group(1)}/{m.group(2)}  {cur}")
