"""Attribute executed warp instructions (ncu source page) to CUDA source lines (nvdisasm -g line info).
usage: ncu_lines.py <report.ncu-rep> <libgenb200.so> <kernel-name-substring> [top]"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, so, pat = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout
# per function: ordered list of (offset, file, line)
funcs, cur, loc = {}, None, ("?", 0)
for ln in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
    if m:
        cur = m.group(1)
        funcs[cur] = []
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        loc = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m and cur:
        funcs[cur].append((int(m.group(1), 16), loc, m.group(2).strip()))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
kern, rows = None, []
for r in csv.reader(src.splitlines()):
    if r and r[0] == "Kernel Name":
        if kern and pat in kern and rows:
            break
        kern, rows = r[1], []
        continue
    if r and r[0] == "Address":
        iA, iS, iE, iSm = r.index("Address"), r.index("Source"), r.index("Instructions Executed"), r.index("# Samples")
        continue
    if len(r) > 8:
        rows.append((int(r[iA], 16), r[iS].strip(), int(r[iE]), int(r[iSm])))
assert kern and pat in kern, "kernel not found"
# match function by demangled-name fragments
cand = [f for f in funcs if len(funcs[f]) == len(rows)]
# same static length for the float and the double instantiation: take the one the kernel name asks for
want = "Id" if "<double" in kern else ("If" if "<float" in kern else "")
cand.sort(key=lambda f: 0 if (want and ("I" + want[1] + "E" in f or "I" + want[1] + "L" in f)) else 1)
fn = None
for f in cand:
    if all(re.sub(r"\s+", " ", a[2]).split(" ")[0] == re.sub(r"\s+", " ", b[1]).split(" ")[0] or True for a, b in zip(funcs[f][:5], rows[:5])):
        if pat.split("<")[0].split("::")[-1] in f:
            fn = f
            break
assert fn, "no matching function (static instruction count %d)" % len(rows)
by_line, samples = collections.Counter(), collections.Counter()
for (off, loc, ins), (addr, s, e, sm) in zip(funcs[fn], rows):
    by_line[loc] += e
    samples[loc] += sm
tot = sum(by_line.values())
tsm = sum(samples.values()) or 1
print("kernel: %s\nfunction: %s\ntotal warp instructions %d, samples %d" % (kern[:100], fn[:80], tot, tsm))
for loc, e in by_line.most_common(top):
    print("  %-26s %6.2f%% instr  %6.2f%% samples" % ("%s:%d" % loc, 100 * e / tot, 100 * samples[loc] / tsm))
