"""krepp_b200/csrc/fixed5.h (the rounding the device applies to the distances of the compact `dist` rows) against printf:
the integer must be exactly the digits "%.5f" prints, including values on and next to rounding boundaries and exact ties."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "native", "fixed5_host.cpp")
OUT = os.path.join(ROOT, "oracle", "_build", "libfixed5_host.so")


@pytest.fixture(scope="module")
def F():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", OUT, SRC], check=True)
    L = C.CDLL(OUT)
    L.fixed5_many.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
    return L


def units(F, vals):
    v = np.ascontiguousarray(vals, dtype=np.float64)
    out = np.zeros(len(v), np.uint32)
    F.fixed5_many(v.ctypes.data, out.ctypes.data, len(v))
    return out


def test_fixed5_units_equal_printf(F):
    rng = np.random.default_rng(11)
    half = (np.arange(0, 60000) * 2 + 1) / 200000.0  # the doubles nearest to every rounding boundary in [0, 0.6) ...
    vals = np.concatenate([
        rng.uniform(0, 0.5, 200000), rng.uniform(0, 2e-4, 20000), 10.0 ** rng.uniform(-10, -0.3, 50000),
        half, np.nextafter(half, 1.0), np.nextafter(half, 0.0),                       # ... and their neighbours
        np.arange(1, 64, 2) / 64.0, np.arange(1, 4096, 2) / 4096.0,                     # exact ties: j / 64 = (3125 j) / 200000
        np.arange(0, 50001) / 100000.0,
        [0.0, 1e-10, 1.2625413546520176e-05, 0.5, 0.499995, 0.000005, 0.999995, 1.0, 3.999995, 39999.999994]])
    got = units(F, vals)
    want = np.array([int(("%.5f" % v).replace(".", "")) for v in vals], dtype=np.uint64)
    bad = np.nonzero(got != want)[0]
    assert len(bad) == 0, [(float(vals[i]).hex(), int(got[i]), int(want[i])) for i in bad[:5]]
    assert "%.5f" % (1 / 64) == "0.01562" and units(F, [1 / 64])[0] == 1562   # tie -> even
    assert "%.5f" % (3 / 64) == "0.04688" and units(F, [3 / 64])[0] == 4688


def test_fixed5_out_of_range(F):
    got = units(F, [-1e-9, float("nan"), 1.7976931348623157e308, float("inf"), 40000.0])
    assert (got == 0xFFFFFFFF).all()
