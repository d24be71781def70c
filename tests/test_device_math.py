"""Accuracy of the device elementary functions (csrc/lh_math.cuh), checked on the CPU.

The header compiles in a host-emulation mode (same algorithm, same coefficients, MUFU seeds
emulated at 20 bits), so the ulp error of lh_log2 / lh_exp2 / lh_exp2m1 / lh_sqrt / lh_rsqrt / lh_div is measured
here against mpmath at 40 digits.  The GPU build of the same source is covered by
tests/test_gpu_parity.py::test_diagnostics_match_oracle."""
import ctypes as C
import os
import subprocess

import mpmath as mp
import numpy as np
import pytest

import workloads as w

ROOT = w.ROOT
CSRC = os.path.join(ROOT, "landhydrology.jl_b200", "csrc")
SRC = os.path.join(ROOT, "tests", "support", "device_math_host.cpp")


@pytest.fixture(scope="module")
def mathlib(tmp_path_factory):
    out = tmp_path_factory.mktemp("lhm") / "liblhm.so"
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-I", CSRC, SRC, "-o", str(out)],
                   check=True, env=env)
    return C.CDLL(str(out))


def _call(lib, name, *arrays):
    n = len(arrays[0])
    y = np.empty(n)
    dp = C.POINTER(C.c_double)
    fn = getattr(lib, name)
    fn.restype = None
    fn(*[np.ascontiguousarray(a).ctypes.data_as(dp) for a in arrays], y.ctypes.data_as(dp), C.c_long(n))
    return y


def _ulp_err(y, exact):
    mp.mp.dps = 40
    errs = []
    for a, e in zip(y, exact):
        e = mp.mpf(e)
        if e == 0:
            errs.append(0.0 if a == 0 else np.inf)
            continue
        ulp = mp.mpf(2) ** (mp.floor(mp.log(abs(e), 2)) - 52)
        errs.append(float(abs(mp.mpf(float(a)) - e) / ulp))
    return np.array(errs)


def test_log2(mathlib):
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(1e-16, 2.0, 4000), 1.0 + rng.uniform(-1e-3, 1e-3, 1000),
                        np.exp(rng.uniform(-40, 40, 2000)), [0.5, 1.0, 2.0, 0.70710678118654752, 1.4142135623730951]])
    y = _call(mathlib, "lhm_log2", x)
    mp.mp.dps = 40
    err = _ulp_err(y, [mp.log(mp.mpf(float(v)), 2) for v in x])
    assert err.max() <= 2.0, err.max()
    sp = _call(mathlib, "lhm_log2", np.array([0.0, -1.0, np.inf, np.nan, -0.5, -0.0]))
    assert -1024 < sp[0] < -1022 and np.isnan(sp[1]) and np.isnan(sp[2]) and np.isnan(sp[3]) and np.isnan(sp[4])
    assert np.isnan(sp[5])      # -0.0 is flagged like any negative number: the closures form 1 - 2^u with
                                # lh_one_minus_exp2, whose zero is +0
    z = _call(mathlib, "lhm_one_minus_exp2", np.array([0.0, -0.0, -1.0, 1.0]))
    assert z[0] == 0.0 and not np.signbit(z[0]) and not np.signbit(z[1]) and z[2] == 0.5 and z[3] == -1.0


def test_exp2_and_exp2m1(mathlib):
    rng = np.random.default_rng(2)
    x = np.concatenate([rng.uniform(-50, 50, 4000), rng.uniform(-1, 1, 3000), rng.uniform(-0.1, 0.1, 1000),
                        rng.uniform(-1000, 1000, 1000), [0.0, -0.0, 1e-300, -1e-20]])
    mp.mp.dps = 40
    y = _call(mathlib, "lhm_exp2", x)
    err = _ulp_err(y, [mp.mpf(2) ** mp.mpf(float(v)) for v in x])
    assert err.max() <= 1.5, err.max()
    ym = _call(mathlib, "lhm_exp2m1", x)
    exact = [mp.expm1(mp.mpf(float(v)) * mp.log(2)) for v in x]
    errm = _ulp_err(ym, exact)
    rel = np.abs(ym - np.array([float(e) for e in exact])) / np.maximum(np.abs(ym), 1e-300)
    assert rel.max() <= 6e-15, rel.max()          # s - 1 carries the table entry's rounding when k != 0
    small = np.abs(x) <= 1 / 128
    assert errm[small].max() <= 8.0               # the cancellation-sensitive range: k == 0, the result is p = r g(r)
                                                  # itself; the degree-4 fit of g is good to 4.5e-16 relative + Horner
                                                  # rounding.  Beyond it s - 1 cancels like the reference's own literal
                                                  # 1 - x^c does (1.1e-16 / |1 - 2^x|): covered by `rel` above.
    sp = _call(mathlib, "lhm_exp2", np.array([-np.inf, -1100.0, -1023.0, np.nan]))
    assert 0 <= sp[0] < 1e-300 and 0 <= sp[1] < 1e-300 and 0 <= sp[2] < 1e-300 and np.isnan(sp[3])
    sp = _call(mathlib, "lhm_exp2m1", np.array([-np.inf, -1100.0, 0.0, np.nan]))
    assert sp[0] == -1 and sp[1] == -1 and sp[2] == 0 and np.isnan(sp[3])


def test_sqrt_rcp_div(mathlib):
    rng = np.random.default_rng(3)
    x = np.concatenate([rng.uniform(1e-16, 4.0, 4000), np.exp(rng.uniform(-200, 200, 2000))])
    mp.mp.dps = 40
    assert _ulp_err(_call(mathlib, "lhm_sqrt", x), [mp.sqrt(mp.mpf(float(v))) for v in x]).max() <= 1.0
    assert _call(mathlib, "lhm_sqrt", np.array([0.0]))[0] == 0.0
    assert np.isnan(_call(mathlib, "lhm_sqrt", np.array([-1.0]))[0])
    assert _ulp_err(_call(mathlib, "lhm_rsqrt", x), [1 / mp.sqrt(mp.mpf(float(v))) for v in x]).max() <= 2.0
    assert _ulp_err(_call(mathlib, "lhm_rcp", x), [1 / mp.mpf(float(v)) for v in x]).max() <= 1.5
    a = rng.uniform(-1e8, 1e8, len(x))
    q = _call(mathlib, "lhm_div", a, x)
    assert _ulp_err(q, [mp.mpf(float(u)) / mp.mpf(float(v)) for u, v in zip(a, x)]).max() <= 1.0


@pytest.mark.parametrize("c", [1.0 / (1.0 - 1.0 / 1.56), 1.0 - 1.0 / 1.56, 1.0 / (1.0 - 1.0 / 3.96), 1.0 - 1.0 / 3.96,
                               0.5, 2.0, (1.0 - 0.24 * 0.92) / 2.0, 0.3896, 3.7])
def test_pow_fixed(mathlib, c):
    """lh_pow_fixed: x^c with per-exponent tables (the van Genuchten 1/m, m and the Kersten exponent)."""
    rng = np.random.default_rng(5)
    x = np.concatenate([rng.uniform(1e-6, 1.4, 6000), 1.0 + rng.uniform(-4e-3, 4e-3, 1500), 2.0 ** rng.uniform(-60, 0, 1500),
                        [0.5, 1.0, 0.70710678118654752, 1.4142135623730951, 1.0 - 2.0 ** -53, 1.0 + 2.0 ** -52]])
    mp.mp.dps = 40
    y = _call(mathlib, "lhm_pow", x, np.array([c, 0.0] + [0.0] * (len(x) - 2)))
    exact = [mp.mpf(float(v)) ** mp.mpf(c) for v in x]
    err = _ulp_err(y, exact)
    assert err.max() <= 2.5, err.max()
    # x^c - 1 and 1 - x^c keep their relative accuracy around x = 1 (central interval: s == 1 exactly)
    near = (x >= 1.0 - 2.0 ** -9) & (x <= 1.0 + 2.0 ** -8)      # the interval around m = 1: r = A = B = 1 exactly
    ym1 = _call(mathlib, "lhm_pow", x, np.array([c, 1.0] + [0.0] * (len(x) - 2)))
    om = _call(mathlib, "lhm_pow", x, np.array([c, 2.0] + [0.0] * (len(x) - 2)))
    ex1 = [e - 1 for e in exact]
    e1 = _ulp_err(ym1[near], [e for e, k in zip(ex1, near) if k])
    assert e1.max() <= 64.0, e1.max()     # degree-4 g: |binom(c, 6)| a^5 / 16 / c <= 7e-15 relative (the literal 1 - x^c of
                                          # the reference is off by 1.1e-16 / |1 - x^c| there: >= 3e-14)
    assert np.array_equal(om, -ym1) or np.allclose(om, -ym1, rtol=1e-15, atol=0)
    rel = np.abs(ym1 - np.array([float(e) for e in ex1])) / np.maximum(np.abs(ym1), 1e-300)
    assert rel[~near & (x > 0.01)].max() <= 3e-13 / min(c, 1.0)
    # below 2^-64 the power reads as 0; negative arguments give finite garbage (callers that need a NaN flag test the
    # sign themselves: lh_pow_bad)
    sp = _call(mathlib, "lhm_pow", np.array([2.0 ** -70, 0.0, -0.3]), np.array([c, 0.0, 0.0]))
    assert sp[0] == 0.0 and sp[1] == 0.0 and np.isfinite(sp[2])
