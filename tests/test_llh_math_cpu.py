"""CPU checks of krepp_b200/csrc/llh_math.cuh (the floating-point core of the solve / place kernels), built for the
host with -ffp-contract=off and compared with the oracle's restatement of HDistHistLLH + Brent
(oracle/krepp_oracle.c: ko_llh / ko_brent, ref src/hdhistllh.hpp:51-96, minima.hpp:23-138)."""
import ctypes as C
import math
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "native", "llh_math_host.cpp")
OUT = os.path.join(ROOT, "oracle", "_build", "libllh_math_host.so")


@pytest.fixture(scope="module")
def M():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", OUT, SRC], check=True)
    L = C.CDLL(OUT)
    L.llh_powi.restype = C.c_double
    L.llh_powi.argtypes = [C.c_double, C.c_uint32]
    L.llh_eval.restype = C.c_double
    L.llh_eval.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_double), C.c_double, C.c_double, C.c_double]
    L.llh_brent.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_double), C.c_double, C.c_double, C.c_int,
                            C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_uint32)]
    L.llh_tables_out.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_double)]
    return L


def _cases(n, seed, th=4):
    rng = np.random.default_rng(seed)
    for _ in range(n):
        hist = np.zeros(17)
        total = int(rng.integers(1, 60))
        for _ in range(total):
            hist[min(th, int(rng.geometric(0.45)) - 1)] += 1
        if rng.random() < 0.2:  # sparse single-hit histograms, the bulk of the 1,000-genome workload
            hist[:] = 0
            hist[int(rng.integers(0, th + 1))] = 1
        on = int(rng.integers(max(total, 20), 125))
        yield hist, float(on - min(total, on)), float(rng.uniform(0.02, 0.6))


def test_tables_match_reference_tables(M):
    for (h, k, th) in [(11, 27, 4), (7, 21, 3), (11, 27, 0), (13, 31, 7)]:
        ck, hnk = (C.c_uint64 * 33)(), (C.c_uint64 * 33)()
        O.lib().ko_llh_tables(h, k, th, ck, hnk)
        w = (C.c_double * 33)()
        M.llh_tables_out(h, k, th, w)
        for x in range(k + 1):
            assert w[x] == float(hnk[x] if x <= th else ck[x]), (h, k, th, x)


def test_powi_is_the_correctly_rounded_power(M):
    from fractions import Fraction
    rng = np.random.default_rng(3)
    xs = np.concatenate([1.0 - rng.uniform(1e-10, 0.5, 4000), rng.uniform(0.5, 1.0, 1000)])
    for n in (1, 2, 21, 27, 31, 32):
        for x in xs[:1500]:
            got = M.llh_powi(float(x), n)
            exact = Fraction(float(x)) ** n
            # correctly rounded <=> |got - exact| <= half an ulp of got
            assert abs(Fraction(got) - exact) <= Fraction(math.ulp(got)) / 2, (x, n, got)


def test_objective_equals_oracle(M):
    """Same operation order as the reference; pow/log may differ from glibc's in the last bit."""
    L = O.lib()
    worst = 0.0
    ident = tot = 0
    for hist, uc, rho in _cases(400, 11):
        hh = (C.c_double * 17)(*hist)
        for d in (1e-10, 1e-4, 0.013, 0.1, 0.3090169943, 0.5):
            a = M.llh_eval(11, 27, 4, hh, uc, rho, d)
            b = L.ko_llh(11, 27, 4, hh, uc, rho, d)
            worst = max(worst, abs(a - b) / max(abs(b), 1e-300))
            ident += a == b
            tot += 1
    assert worst < 1e-14, worst
    assert ident / tot > 0.9, (ident, tot)


@pytest.mark.parametrize("memo", [0, 1])
def test_brent_equals_oracle(M, memo):
    L = O.lib()
    worst = 0.0
    ident = tot = 0
    for hist, uc, rho in _cases(3000, 5):
        hh = (C.c_double * 17)(*hist)
        d, v, hits = C.c_double(), C.c_double(), C.c_uint32()
        M.llh_brent(11, 27, 4, hh, uc, rho, memo, C.byref(d), C.byref(v), C.byref(hits))
        rd, rv, it = C.c_double(), C.c_double(), C.c_uint32()
        L.ko_brent(11, 27, 4, hh, uc, rho, C.byref(rd), C.byref(rv), C.byref(it))
        worst = max(worst, abs(d.value - rd.value) / rd.value)
        ident += d.value == rd.value
        tot += 1
        if memo:
            assert hits.value == 3  # evaluations 0, 1, 2 are always served by the table
    assert worst < 1e-9, worst       # same iteration path: differences are last-bit effects of pow/log only
    assert ident / tot > 0.9, (ident, tot)
