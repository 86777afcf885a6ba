"""Registers / spills per kernel from a `python gen.jl_b200/build.py --force -v` log (ptxas -v).
usage: python tools/ptxas_report.py build.log [substring]"""
import re
import subprocess
import sys

log = open(sys.argv[1]).read().split("\n")
pat = sys.argv[2] if len(sys.argv) > 2 else ""
rows, name, spill = [], None, ""
for line in log:
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        name, spill = m.group(1), ""
    if "spill" in line:
        spill = line.strip()
    m = re.search(r"Used (\d+) registers", line)
    if m and name:
        rows.append((name, int(m.group(1)), spill))
        name = None
dem = subprocess.run(["c++filt"] + [r[0] for r in rows], capture_output=True, text=True).stdout.split("\n")
for (n, r, sp), d in zip(rows, dem):
    if pat in d:
        print("%3d regs | %s | %s" % (r, sp.replace(" bytes", "B"), d[:160]))
