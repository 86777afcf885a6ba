"""Host-side timeline of the sharded maybe_resample_ (run under torchrun on a multi-GPU box)."""
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as entry  # noqa: E402
import fixtures as F  # noqa: E402

lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
g = entry.load_package()
from genb200 import sharded, _lib as L  # noqa: E402
L.check(L.lib().gb_set_device(lr))
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
zs = F.bearings_series(40)
spf = sharded.ShardedParticleFilter(g.bearings_model(), n, seed=1).initialize(zs[0])
if len(sys.argv) > 2 and sys.argv[2] == "async_first":
    spf.run_([zs[t] for t in range(1, 4)], ess_threshold=n + 1)
    for t in range(4, 7):
        spf.particle_filter_step_((t,), zs[t])
marks = {}


def wrap(obj, name):
    f = getattr(obj, name)

    def w(*a, **k):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = f(*a, **k)
        torch.cuda.synchronize()
        marks[name] = marks.get(name, 0.0) + time.perf_counter() - t0
        return r
    setattr(obj, name, w)


for nm in ("weight_max_device", "weight_sums_device", "systematic_offspring", "export_offspring", "import_ext", "commit_lazy",
           "offspring_local"):
    wrap(spf.ops, nm)
wrap(spf, "_all_to_all_v")
wrap(spf, "_global_stats")
wrap(sharded, "plan_exchange")
for t in range(1, 6):
    spf.maybe_resample_(ess_threshold=n + 1)
    spf.particle_filter_step_((t,), zs[t])
marks.clear()
torch.cuda.synchronize()
dist.barrier()
T0 = time.perf_counter()
K = 10
tr = ts = 0.0
for t in range(6, 6 + K):
    torch.cuda.synchronize(); a = time.perf_counter()
    spf.maybe_resample_(ess_threshold=n + 1)
    torch.cuda.synchronize(); b = time.perf_counter()
    spf.particle_filter_step_((t,), zs[t])
    torch.cuda.synchronize(); c = time.perf_counter()
    tr += b - a
    ts += c - b
tot = time.perf_counter() - T0
if dist.get_rank() == 0:
    print("per step: total %.3f ms  resample %.3f ms  step %.3f ms" % (1e3 * tot / K, 1e3 * tr / K, 1e3 * ts / K))
    for k, v in sorted(marks.items(), key=lambda kv: -kv[1]):
        print("   %-24s %.3f ms/step" % (k, 1e3 * v / K))
dist.destroy_process_group()
