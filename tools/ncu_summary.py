"""Summarise an .ncu-rep: per-kernel headline metrics + executed-instruction mix (for profiles/*.md)."""
import collections
import csv
import re
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keys = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
        "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_wait_per_warp_active.pct", "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct"]
for r in rows[2:]:
    print("## " + r[hdr.index("Kernel Name")][:110])
    for k in keys[1:]:
        if k in hdr:
            i = hdr.index(k)
            print("  %-75s %s %s" % (k, r[i][:30], units[i]))
    # warp stall reasons (warps stalled per issue slot), largest first
    pre, suf = "smsp__average_warps_issue_stalled_", "_per_issue_active.ratio"
    st = sorted(((float(r[i]), hdr[i][len(pre):-len(suf)]) for i in range(len(hdr)) if hdr[i].startswith(pre) and r[i]), reverse=True)
    if st:
        print("  stalls (warps per issue-active cycle): " + ", ".join("%s %.2f" % (n, v) for v, n in st[:7]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
cur = None
tally = {}
for r in csv.reader(src.splitlines()):
    if r and r[0] == "Kernel Name":
        cur = r[1][:90]
        tally.setdefault(cur, collections.Counter())
        continue
    if not r or r[0] == "Address" or len(r) < 8:
        if r and r[0] == "Address":
            iS, iE = r.index("Source"), r.index("Instructions Executed")
        continue
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[iS].strip())
    if m:
        tally[cur][m.group(2)] += int(r[iE])
for k, c in tally.items():
    tot = sum(c.values())
    print("## instruction mix (warp instructions executed): " + k)
    print("  total %d" % tot)
    print("  " + ", ".join("%s %.1f%%" % (op, 100 * v / tot) for op, v in c.most_common(14)))
