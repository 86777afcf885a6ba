"""CPU-only checks of the C-ABI library: it loads, exports every symbol include/genb200.h declares,
fails loudly without a GPU (no CPU fallback), and its exact-integer helpers (the host-side instances
of the device functions) agree with Python big integers and with the oracle."""
import ctypes as C
import os
import random
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "genb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(g):
    from genb200 import _lib as L
    lib = C.CDLL(g.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 50
    for name in names:
        assert hasattr(lib, name), "libgenb200.so does not export %s" % name
        assert name in L.SIGNATURES, "ctypes binding is missing %s" % name
    assert sorted(L.SIGNATURES) == names


def test_no_cpu_fallback(g):
    """Without a GPU every compute entry must fail with GB_ENOGPU -- never compute on the host."""
    if g.device_count() > 0:
        pytest.skip("GPU present")
    from genb200 import _lib as L
    m = g.lgssm_model()
    with pytest.raises(g.GenB200Error) as ei:
        g.initialize_particle_filter(m, (0,), 0.3, 100)
    assert ei.value.code == L.ENOGPU
    with pytest.raises(g.GenB200Error):
        g.resample(np.zeros(10))
    with pytest.raises(g.GenB200Error):
        g.logpdf("normal", [0.0], [0.0], [1.0])
    with pytest.raises(g.GenB200Error):
        g.importance_sampling(m, (0,), 0.3, 10)


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "gen.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                for line in open(os.path.join(dirpath, f)):
                    if re.search(r"#\s*include|\bimport\b|CDLL|dlopen|sys\.path", line):
                        assert "oracle" not in line.lower(), (f, line)


def test_model_validation(g):
    from genb200 import _lib as L
    with pytest.raises(g.GenB200Error) as ei:
        g.DeviceModel(99, [1.0], "bogus")
    assert ei.value.code == L.EUNSUPPORTED
    with pytest.raises(g.GenB200Error):
        g.DeviceModel(L.MODEL_LGSSM, [1.0, 2.0], "short")
    m = g.bearings_model()
    assert m.state_dim == 4 and m.num_obs == 1 and m.num_normals(True) == 4 and m.num_normals(False) == 2
    h = g.hmm_model([0.5, 0.5], np.eye(2), np.eye(2))
    assert h.state_dim == 1 and h.num_uniforms(False) == 1
    # argument checks come before any device work: maximum sizes and shard geometry are rejected with GB_EINVAL
    for n_local, n_global, off in ((2 ** 31, 2 ** 31, 0), (0, 10, 0), (10, 5, 0), (6, 10, 5)):
        with pytest.raises(g.GenB200Error) as ei:
            g.create_particle_filter(m, n_local, n_global=n_global, pid_offset=off)
        assert ei.value.code == L.EINVAL
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(g.GenB200Error) as ei:
            g.create_particle_filter(m, 10, dtype="f64")
        assert ei.value.code == L.ENOGPU                                 # ... and everything else needs the GPU


def test_philox_matches_oracle_and_known_answer(g, O):
    from genb200 import _lib as L
    lib = L.lib()
    # Random123 known-answer vectors for philox4x32-10
    kats = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
            ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
            ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
             (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, exp in kats:
        c = (C.c_uint32 * 4)(*ctr)
        k = (C.c_uint32 * 2)(*key)
        o = (C.c_uint32 * 4)()
        lib.gb_selftest_philox(c, k, o)
        assert tuple(o) == exp
        assert tuple(O.philox(ctr, key)) == exp
    rnd = random.Random(1)
    for _ in range(200):
        seed, step = rnd.getrandbits(64), rnd.getrandbits(32)
        assert lib.gb_philox_resample_u0(seed, step) == O.lib().go_philox_resample_u0(seed, step)
    a = g.philox_multinomial_u64(12345, 7, 1000, 257)
    assert np.array_equal(a, O.philox_multinomial_u64(12345, 7, 1000, 257))


def test_quantize_matches_oracle(g, O):
    from genb200 import _lib as L
    lib = L.lib()
    rng = np.random.default_rng(0)
    xs = np.concatenate([-np.abs(rng.standard_normal(2000)) * 10, [-0.0, 0.0, -1e-300, -707.9, -708.0, -709.0, -1e9,
                                                                    -np.inf, np.nan, 1.0]])
    for b in (31, 35, 42, 52):
        for x in xs:
            assert lib.gb_selftest_quantize(float(x), b) == O.lib().go_quantize(float(x), b), (x, b)
    for n in (1, 2, 3, 1000, 10 ** 6, 10 ** 8, 2 ** 30, 2 ** 31 - 1):
        assert lib.gb_quant_bits(n) == O.lib().go_quant_bits(n)


def test_div128_exact(g):
    from genb200 import _lib as L
    lib = L.lib()
    rnd = random.Random(7)
    for _ in range(20000):
        dbits = rnd.randint(31, 62)
        d = rnd.getrandbits(dbits) | (1 << (dbits - 1))
        qbits = rnd.randint(1, 62)
        q = rnd.getrandbits(qbits)
        r = rnd.randrange(d) if rnd.random() < 0.8 else rnd.choice([0, d - 1])
        a = q * d + r
        assert lib.gb_selftest_div128(a >> 64, a & (2 ** 64 - 1), d) == q, (a, d)


def test_slots_below_matches_bigint_definition(g):
    """#{k : floor((k 2^32 + u0) Q / (N 2^32)) < C} -- the device's closed form vs brute force / big ints."""
    from genb200 import _lib as L
    lib = L.lib()
    rnd = random.Random(3)
    # brute force on small N
    for _ in range(300):
        n = rnd.randint(1, 40)
        Q = rnd.getrandbits(rnd.randint(32, 62)) | (1 << 31)
        u0 = rnd.getrandbits(32)
        taus = [((k << 32) + u0) * Q // (n << 32) for k in range(n)]
        for C_ in [0, 1, Q, Q // 2, Q - 1] + [rnd.randrange(Q + 1) for _ in range(10)] + taus[:5] + [t + 1 for t in taus[:5]]:
            C_ = min(C_, Q)
            assert lib.gb_selftest_slots_below(C_, n, Q, u0) == sum(1 for t in taus if t < C_)
    # closed form with big ints on large N
    for _ in range(5000):
        n = rnd.randint(1, 2 ** 31 - 1)
        lg = max(0, (n - 1).bit_length())
        Q = rnd.randrange(1 << 31, (1 << 62) + 1 if lg == 0 else (n << (62 - lg)) + 1)
        Q = min(Q, 1 << 62)
        u0 = rnd.getrandbits(32)
        C_ = rnd.randrange(Q + 1)
        if C_ == 0:
            exp = 0
        else:
            X = (C_ * (n << 32) - 1) // Q
            exp = 0 if X < u0 else ((X - u0) >> 32) + 1
        assert lib.gb_selftest_slots_below(C_, n, Q, u0) == exp
    # residual sweep: tau_k = floor((k 2^32 + u0) Q / 2^32), C up to R*Q < 2^93
    for _ in range(5000):
        Q = rnd.getrandbits(rnd.randint(32, 62)) | (1 << 31)
        R = rnd.randint(1, 2 ** 31 - 1)
        C_ = rnd.randrange(R * Q + 1)
        u0 = rnd.getrandbits(32)
        if C_ == 0:
            exp = 0
        else:
            X = ((C_ << 32) - 1) // Q
            exp = 0 if X < u0 else ((X - u0) >> 32) + 1
        assert lib.gb_selftest_slots_below_residual(C_ >> 64, C_ & (2 ** 64 - 1), Q, u0) == exp


def test_gbmath_accuracy_against_mpmath(g):
    """The engine's own fp64 log / sincos / atan2 / exp (csrc/gbmath.cuh) use only integer ops, fma and
    IEEE + * / sqrt, so this host instance is bit-identical to the device one: accuracy <= 2 ulp."""
    import math
    import mpmath as mp
    from genb200 import _lib as L
    f = L.lib().gb_selftest_math
    mp.mp.prec = 120
    rnd = random.Random(1)

    def ulps(got, exact):
        e = float(exact)
        return abs(mp.mpf(got) - exact) / (math.ulp(e) if e != 0 else mp.mpf(2) ** -1074)

    worst = {"log": 0, "sin": 0, "cos": 0, "atan2": 0, "exp": 0}
    for _ in range(4000):
        for x in (rnd.random(), 2.0 ** (-rnd.uniform(0, 53)), 1 - 2.0 ** (-rnd.uniform(1, 52)), rnd.uniform(0.5, 2.0),
                  10 ** rnd.uniform(-300, 300)):
            worst["log"] = max(worst["log"], ulps(f(0, x, 0), mp.log(mp.mpf(x))))
        u = rnd.random()
        worst["sin"] = max(worst["sin"], abs(mp.mpf(f(1, u, 0)) - mp.sin(2 * mp.pi * mp.mpf(u))) * 2 ** 53)
        worst["cos"] = max(worst["cos"], abs(mp.mpf(f(2, u, 0)) - mp.cos(2 * mp.pi * mp.mpf(u))) * 2 ** 53)
        y, x = rnd.gauss(0, 1) * 10 ** rnd.uniform(-3, 3), rnd.gauss(0, 1) * 10 ** rnd.uniform(-3, 3)
        worst["atan2"] = max(worst["atan2"], ulps(f(3, y, x), mp.atan2(mp.mpf(y), mp.mpf(x))))
        y, x = rnd.uniform(0.9, 1.0), rnd.uniform(-0.05, 0.05)      # the bearings-only regime
        worst["atan2"] = max(worst["atan2"], ulps(f(3, y, x), mp.atan2(mp.mpf(y), mp.mpf(x))))
        a = rnd.uniform(-40, 40)
        worst["exp"] = max(worst["exp"], ulps(f(4, a, 0), mp.exp(mp.mpf(a))))
    assert all(v < 2.0 for v in worst.values()), worst
    # special values / quadrants
    assert f(1, 0.25, 0) == 1.0 and f(2, 0.5, 0) == -1.0 and f(1, 0.75, 0) == -1.0 and f(2, 0.0, 0) == 1.0
    assert f(3, 0.0, -1.0) == math.pi and f(3, 1.0, 0.0) == math.pi / 2 and f(3, -1.0, 0.0) == -math.pi / 2
    assert f(3, 0.0, 0.0) == 0.0 and f(0, 1.0, 0) == 0.0
    for _ in range(1000):                     # the bit-trick uniform equals the reference definition
        hi, lo = rnd.getrandbits(32), rnd.getrandbits(32)
        v = ((hi << 32) | lo) >> 12
        assert f(5, float(hi), float(lo)) == (v + 0.5) * 2.0 ** -52


def test_spec_programs_lower_to_cuda_and_compile_for_sm100a_without_a_gpu(g):
    """GB_MODEL_SPEC run-time specialisation: the generated translation unit has one statement per node and NVRTC
    compiles it for sm_100a here (no GPU needed to compile; loading and launching are covered by the gpu tests)."""
    gen, step = g.SpecProgram(), g.SpecProgram()
    x = gen.sample_normal(gen.param(0), gen.param(1))
    gen.observe_normal(gen.obs(0), x, gen.const(0.7))
    gen.store(0, x)
    x = step.sample_normal(step.mul(step.const(0.9), step.prev(0)), step.param(1))
    b = step.sample_bernoulli(step.const(0.25))
    step.observe_gamma(step.obs(0), step.const(2.0), step.exp(x))
    step.store(0, step.add(x, b))
    model = g.spec_model(1, [0.0, 1.5], 1, gen, step)
    src = g.spec_jit_source(model)
    assert "spec_jit_kernel" in src and "philox_normal_pair" in src and "logpdf_gamma<R>" in src
    assert "(R)(0x1.8p+0)" in src                       # parameters are baked in as exact literals
    assert src.count("st_stream(sout") == 2
    for dtype in ("f64", "f32"):
        assert g.spec_jit_compile(model, dtype) > 10000   # a cubin came back
    with pytest.raises(g.GenB200Error):
        g.spec_jit_compile(g.lgssm_model(), "f64")        # presets are compiled ahead of time


def test_spec_jit_disk_cache(g, tmp_path, monkeypatch):
    """GENB200_JIT_CACHE=<dir>: the compiled program is stored under a hash of (source, headers, dtype) and reused."""
    import time
    monkeypatch.setenv("GENB200_JIT_CACHE", str(tmp_path))

    def model(c):
        gen = g.SpecProgram()
        x = gen.sample_normal(gen.const(c), gen.const(1.0))
        gen.observe_normal(gen.obs(0), x, gen.const(0.5))
        gen.store(0, x)
        return g.spec_model(1, [], 1, gen)
    t0 = time.perf_counter()
    nb = g.spec_jit_compile(model(0.25), "f64")
    t_cold = time.perf_counter() - t0
    files = sorted(p.name for p in tmp_path.iterdir())
    assert len(files) == 1 and files[0].startswith("genb200_") and files[0].endswith(".jit")
    t0 = time.perf_counter()
    assert g.spec_jit_compile(model(0.25), "f64") == nb            # a new model handle with the same program: cache hit
    assert time.perf_counter() - t0 < t_cold / 5
    g.spec_jit_compile(model(0.5), "f64")                           # different literal -> different program
    g.spec_jit_compile(model(0.25), "f32")                          # different dtype -> different program
    assert len(list(tmp_path.iterdir())) == 3
