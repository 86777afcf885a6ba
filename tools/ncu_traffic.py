"""DRAM traffic of the dominant kernel from an `ncu --set full` capture -> profiles/r2_step_kernel_traffic.json, the file
bench.py reads for `roofline.traffic` (measured bytes, not a constant in the source).
usage: python tools/ncu_traffic.py <report.ncu-rep> <particles> [kernel-substring] [out.json]"""
import csv
import json
import os
import subprocess
import sys

rep, n = sys.argv[1], int(float(sys.argv[2]))
pat = sys.argv[3] if len(sys.argv) > 3 else "step_kernel<double, 3, 0, 0, 0, 1>"
out = sys.argv[4] if len(sys.argv) > 4 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles",
                                                          "r2_step_kernel_traffic.json")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
iname, ird, iwr, itime = (hdr.index(k) for k in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"))
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
hits = [r for r in rows[2:] if pat in r[iname]]
if not hits:
    sys.exit("no launch of %r in %s" % (pat, rep))
rd = [float(r[ird].replace(",", "")) * scale[units[ird]] for r in hits]
wr = [float(r[iwr].replace(",", "")) * scale[units[iwr]] for r in hits]
doc = {"file": "profiles/" + os.path.basename(rep).replace(".ncu-rep", "_kernels_ncu.txt"), "report": os.path.basename(rep),
       "kernel": hits[0][iname], "launches": len(hits), "particles": n, "dtype": "f64" if "<double" in pat else "f32", "state_dim": 4,
       "dram_bytes_read_per_launch": sum(rd) / len(rd), "dram_bytes_write_per_launch": sum(wr) / len(wr),
       "dram_bytes_per_particle": (sum(rd) + sum(wr)) / len(hits) / n,
       "gpu_time_ms_under_ncu": [float(r[itime].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6}.get(units[itime], 1.0) for r in hits]}
json.dump(doc, open(out, "w"), indent=1)
print(json.dumps(doc))
