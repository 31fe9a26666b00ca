"""Accuracy of the shared deterministic elementary functions (galacticus_b200/csrc/glc_detmath.h)
against numpy/glibc; their bit-reproducibility GPU-vs-CPU is what the gpu parity tests assert."""
import ctypes as C

import numpy as np
import pytest


@pytest.fixture(scope="module")
def dm(oracle_lib):
    L = oracle_lib.lib()
    dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
    L.orc_dm_eval.argtypes = [C.c_int, C.c_long, dp, dp, dp]

    def run(which, x, y=None):
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = np.zeros_like(x)
        L.orc_dm_eval(which, x.size, x, x if y is None else np.ascontiguousarray(y, dtype=np.float64), out)
        return out

    return run


def ulp(a, b):
    return np.abs(a - b) / np.spacing(np.abs(b))


def test_exp_log_atan_cbrt(dm):
    rng = np.random.default_rng(1)
    n = 400_000
    x = rng.uniform(-700, 700, n)
    assert ulp(dm(0, x), np.exp(x)).max() <= 2
    x = 10.0 ** rng.uniform(-300, 300, n)
    assert ulp(dm(1, x), np.log(x)).max() <= 3
    x = rng.uniform(0.5, 2.0, n)
    assert np.abs(dm(1, x) - np.log(x)).max() <= 2.3e-16
    x = 10.0 ** rng.uniform(-8, 8, n) * rng.choice([-1.0, 1.0], n)
    assert ulp(dm(3, x), np.arctan(x)).max() <= 2
    x = 10.0 ** rng.uniform(-200, 200, n)
    assert ulp(dm(4, x), np.cbrt(x)).max() <= 2


def test_pow(dm):
    rng = np.random.default_rng(2)
    n = 400_000
    x = 10.0 ** rng.uniform(-10, 15, n)
    y = rng.uniform(-4, 4, n)
    r = dm(2, x, y)
    assert (np.abs(r - np.power(x, y)) / np.power(x, y)).max() < 1e-13
    assert dm(2, np.array([2.0, 9.0, 1.0, 5.0]), np.array([2.0, 0.5, 7.3, 0.0])).tolist() == [4.0, 3.0, 1.0, 1.0]


def test_special_values(dm):
    assert np.isinf(dm(0, np.array([1000.0]))[0]) and dm(0, np.array([-1000.0]))[0] == 0.0
    assert dm(1, np.array([1.0]))[0] == 0.0 and np.isneginf(dm(1, np.array([0.0]))[0])
    assert np.isnan(dm(1, np.array([-1.0]))[0])
    assert dm(4, np.array([0.0, -8.0])).tolist() == [0.0, -2.0]
