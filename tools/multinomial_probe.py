"""Raw multinomial resampling (reference semantics: N iid categorical draws) at N = 1e8: time per call."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as entry
g = entry.load_package()
import ctypes as C
import torch
from genb200 import _lib as L
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10 ** 8
ld = (n + 2047) // 2048 * 2048
gen = torch.Generator(device="cuda").manual_seed(20244)
lw = torch.zeros(ld, dtype=torch.float64, device="cuda")
lw[:n] = 3.0 * torch.randn(n, dtype=torch.float64, device="cuda", generator=gen)
anc = torch.empty(ld, dtype=torch.int32, device="cuda")
for kind, name in ((L.RESAMPLE_SYSTEMATIC, "systematic"), (L.RESAMPLE_MULTINOMIAL, "multinomial"), (L.RESAMPLE_RESIDUAL, "residual")):
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        L.check(L.lib().gb_resample_device(C.c_void_p(lw.data_ptr()), n, L.F64, kind, 12345, None, 7, 3, C.c_void_p(anc.data_ptr()), None))
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    cnt = torch.bincount(anc[:n].long(), minlength=n)
    w = torch.exp(lw[:n] - lw[:n].max()); w /= w.sum()
    top = torch.argsort(w)[-1000:]
    print("%-12s N=%g: %.3f ms; offspring of the 1000 heaviest: %d (expected %.1f)" % (name, n, 1e3 * dt, int(cnt[top].sum()), float(n * w[top].sum())))
