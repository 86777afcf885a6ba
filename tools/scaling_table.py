"""Scaling table from bench lines: python tools/scaling_table.py <prefix>   (reads <prefix>_{1,2,4,8}gpu.json)"""
import json
import os
import sys

prefix = sys.argv[1]
base = None
for n in (1, 2, 4, 8):
    f = "%s_%dgpu.json" % (prefix, n)
    if not os.path.exists(f):
        continue
    d = json.loads(open(f).read().strip().splitlines()[-1])
    if n == 1:
        base = d
    w = d.get("weak") or {}
    k = {kk.split("(")[0][:9]: round(v["ms_per_step"], 4) for kk, v in d["kernels"].items()}
    par = d.get("sharded_parity")
    col = d.get("collectives")
    print("N=%d %s: %.4f ms/step, %.4g p-steps/s, x%.2f | e2e %.4f ms, %.4g, x%.2f | weak %s ms eff %s e2e %s | %s | parity %s | torch.distributed calls %s" % (
        n, d["scaling"], d["ms_per_step"], d["value"], d["value"] / base["value"], d["e2e"]["ms_per_step"], d["e2e"]["value"],
        d["e2e"]["value"] / base["e2e"]["value"],
        ("%.4f" % w["ms_per_step"]) if w else "-", ("%.3f" % (base["ms_per_step"] / w["ms_per_step"])) if w else "-",
        ("%.4f" % w["e2e_ms_per_step"]) if w else "-", k, par and par["bit_exact"],
        col and (col["torch_distributed_calls_in_timed_region"], col["per_step_in_e2e_region"])))
