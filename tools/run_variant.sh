#!/bin/bash
# usage: tools/run_variant.sh <label> <lib or ""> [bench args]   (run on the GPU box)
LABEL=$1; LIB=$2; shift 2
if [ -n "$LIB" ]; then export GENB200_LIB=$PWD/$LIB; fi
timeout 300 python bench.py --no-cpu "$@" > gpurun_out/bench_$LABEL.json 2> gpurun_out/bench_$LABEL.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_$LABEL.json"))
    print("$LABEL:", "%.4g" % d["value"], "%.3f ms/step" % d["ms_per_step"], {k: round(v["ms_per_step"], 4) for k, v in d["kernels"].items()})
except Exception as e:
    print("$LABEL: no json", e); print(open("gpurun_out/bench_$LABEL.err").read()[-800:])
PY
