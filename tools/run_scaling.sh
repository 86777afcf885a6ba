#!/bin/bash
# usage: tools/run_scaling.sh <ngpus> <scaling> [extra bench args]   (run on the GPU box)
N=$1; SC=$2; shift 2
OUT=gpurun_out/bench_${N}gpu_${SC}.json
if [ "$N" = "1" ]; then
  timeout 600 python bench.py --gpus 1 --scaling $SC "$@" > $OUT 2> gpurun_out/bench_${N}gpu_${SC}.err
else
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --scaling $SC "$@" > $OUT 2> gpurun_out/bench_${N}gpu_${SC}.err
fi
echo "rc=$? ($N gpus, $SC)"
python - <<PY
import json
try:
    d = json.load(open("$OUT"))
    print("$N gpus $SC:", "%.4g" % d["value"], "p-steps/s", "%.3f ms/step" % d["ms_per_step"],
          {k: round(v["ms_per_step"], 4) for k, v in d["kernels"].items()}, "e2e %.4g" % d["e2e"]["value"])
except Exception as e:
    print("no json:", e)
    print(open("gpurun_out/bench_${N}gpu_${SC}.err").read()[-1500:])
PY
