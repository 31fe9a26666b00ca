"""Pin the oracle's nested numerics against the reference's own known-answer tests:
source/tests/integration.F90:40-67 (integrator = gsl_integration_qag) and source/tests/root_finding.F90:56-88
(rootFinder, Brent branch, with and without range expansion).  Same functions, same tolerances."""
import ctypes as C
import math

import pytest

FN1 = C.CFUNCTYPE(C.c_double, C.c_double, C.c_void_p)


class RootFinder(C.Structure):
    _fields_ = [("f", FN1), ("ctx", C.c_void_p), ("tol_abs", C.c_double), ("tol_rel", C.c_double),
                ("expand_type", C.c_int), ("expand_upward", C.c_double), ("expand_downward", C.c_double),
                ("sign_expect_upward", C.c_int), ("sign_expect_downward", C.c_int),
                ("upward_limit_set", C.c_int), ("downward_limit_set", C.c_int),
                ("upward_limit", C.c_double), ("downward_limit", C.c_double), ("n_eval", C.c_int), ("n_iter", C.c_int)]


@pytest.fixture(scope="module")
def L(oracle_lib):
    lib = oracle_lib.lib()
    lib.orc_root_init.argtypes = [C.POINTER(RootFinder), FN1, C.c_void_p, C.c_double, C.c_double]
    lib.orc_root_find.restype = C.c_double
    lib.orc_root_find.argtypes = [C.POINTER(RootFinder), C.c_double, C.c_double, C.c_int, C.c_double, C.c_double,
                                  C.POINTER(C.c_int)]
    lib.orc_qag15.restype = C.c_int
    lib.orc_qag15.argtypes = [FN1, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int,
                              C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int)]
    return lib


def qag(L, f, a, b, epsabs, epsrel):
    cb = FN1(lambda x, _ctx: f(x))
    res, err, n = C.c_double(0), C.c_double(0), C.c_int(0)
    st = L.orc_qag15(cb, None, a, b, epsabs, epsrel, 1000, C.byref(res), C.byref(err), C.byref(n))
    return st, res.value


def test_integration_kats(L):
    st, v = qag(L, lambda x: x, 0.0, 1.0, 0.0, 1.0e-6)  # integration.F90:44-48
    assert st == 0 and abs(v - 0.5) <= 1.0e-6 * 0.5
    st, v = qag(L, math.sin, 0.0, 2.0 * math.pi, 1.0e-6, 0.0)  # :50-54
    assert st == 0 and abs(v) <= 1.0e-6
    st, v = qag(L, lambda x: 1.0 / math.sqrt(x) if x > 0 else 0.0, 0.0, 10.0, 0.0, 1.0e-6)  # :56-60
    assert abs(v - 2.0 * math.sqrt(10.0)) <= 1.0e-6 * 2.0 * math.sqrt(10.0)
    # :62-67 nested: f(x,y) = y cos x, y in 0..x
    inner = lambda x: qag(L, lambda y: y * math.cos(x), 0.0, x, 0.0, 1.0e-6)[1]
    st, v = qag(L, inner, 0.0, 2.0 * math.pi, 0.0, 1.0e-6)
    assert abs(v - 2.0 * math.pi) <= 1.0e-6 * 2.0 * math.pi


def find(L, f, lo, hi, expand=None):
    cb = FN1(lambda x, _ctx: f(x))
    r = RootFinder()
    L.orc_root_init(C.byref(r), cb, None, 1.0e-6, 1.0e-6)
    if expand:
        r.expand_type, r.expand_upward, r.expand_downward = expand
    st = C.c_int(0)
    x = L.orc_root_find(C.byref(r), lo, hi, 0, 0.0, 0.0, C.byref(st))
    return st.value, x


def close(a, b):
    return abs(a - b) <= 1.0e-6 + 1.0e-6 * abs(b)


def test_root_finding_kats(L):
    st, x = find(L, lambda x: x, -1.0, 1.0)  # root_finding.F90:57-61
    assert st == 0 and close(x, 0.0)
    quad = lambda x: x * x - 5.0 * x + 1.0
    st, x = find(L, quad, -1.0, 1.0)  # :63-66
    assert st == 0 and close(x, 0.5 * (5.0 - math.sqrt(21.0)))
    st, x = find(L, quad, 2.0, 10.0)  # :73-76
    assert st == 0 and close(x, 0.5 * (5.0 + math.sqrt(21.0)))
    xexp = lambda x: x * math.exp(-x) + 1.0
    st, x = find(L, xexp, -1.0, 1.0)  # :78-81
    assert st == 0 and close(x, -0.567143)
    # :83-88 root bracketing from a guess: rangeExpand additive +-0.1 around xGuess = 0
    st, x = find(L, xexp, 0.0, 0.0, expand=(1, 0.1, -0.1))
    assert st == 0 and close(x, -0.567143)
