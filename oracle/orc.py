"""ORACLE -- TEST INFRASTRUCTURE ONLY.  ctypes front end to oracle/_build/liborc*.so.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  It is the checker, never the product path.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from galacticus_b200 import abi

HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS: dict[str, C.CDLL] = {}

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def build(force: bool = False) -> None:
    """Compile the oracle with the committed Makefile (gcc)."""
    so = os.path.join(HERE, "_build", "liborc.so")
    if force or not os.path.exists(so) or _stale(so):
        subprocess.run(["make", "-C", HERE, "-s"], check=True)


def _stale(so: str) -> bool:
    t = os.path.getmtime(so)
    srcs = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".c", ".h"))]
    srcs.append(abi.HEADER)
    return any(os.path.getmtime(s) > t for s in srcs)


def lib(fast: bool = False) -> C.CDLL:
    name = "liborc_fast.so" if fast else "liborc.so"
    if name not in _LIBS:
        path = os.path.join(HERE, "_build", name)
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.orc_kat_sin.restype = C.c_double
        L.orc_kat_sin.argtypes = [C.c_double, C.c_double, C.c_double, C.POINTER(C.c_ulong)]
        L.orc_kat_harmonic.restype = None
        L.orc_kat_harmonic.argtypes = [C.c_double, _dp]
        L.orc_params_default.restype = None
        L.orc_params_default.argtypes = [C.POINTER(abi.glc_params), C.c_int]
        L.orc_tables_create.restype = C.c_void_p
        L.orc_tables_destroy.argtypes = [C.c_void_p]
        L.orc_tables_set.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, _dp, C.c_void_p, _dp]
        L.orc_evolve_batch.argtypes = [C.POINTER(abi.glc_params), C.c_void_p, C.c_long, _dp, _ip, _dp,
                                       _ip, _ip, C.POINTER(abi.glc_counters), C.c_int]
        L.orc_rhs_node.argtypes = [C.POINTER(abi.glc_params), C.c_void_p, _dp, C.c_int, _dp,
                                   C.POINTER(C.c_int)]
        L.orc_forest_evolve.argtypes = [C.POINTER(abi.glc_params), C.c_void_p, C.c_long, _ip, _dp, _dp, _dp, _dp, _dp, _ip, _ip,
                                        C.POINTER(abi.glc_forest_counters), C.POINTER(abi.glc_counters), C.c_int]
        L.orc_profiler_reset.restype = None
        L.orc_profiler_reset.argtypes = [C.POINTER(abi.glc_params)]
        L.orc_profiler_read.restype = None
        L.orc_profiler_read.argtypes = [C.POINTER(abi.glc_profile)]
        _LIBS[name] = L
    return _LIBS[name]


def params_default(model: int) -> abi.glc_params:
    p = abi.glc_params()
    lib().orc_params_default(C.byref(p), model)
    return p


class Oracle:
    """CPU node evolver with the calling convention of galacticus_b200.Evolver."""

    def __init__(self, fast: bool = False):
        self.L = lib(fast)
        self.T = C.c_void_p(self.L.orc_tables_create())
        self.params = None

    def __del__(self):
        try:
            self.L.orc_tables_destroy(self.T)
        except Exception:
            pass

    def set_params(self, p: abi.glc_params) -> None:
        self.params = p

    def set_table(self, table_id: int, x0, x1, values) -> None:
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        values = np.ascontiguousarray(values, dtype=np.float64)
        n0 = x0.size
        if x1 is None:
            n1 = values.size // n0
            x1p = None
        else:
            x1 = np.ascontiguousarray(x1, dtype=np.float64)
            n1 = x1.size
            x1p = x1.ctypes.data_as(C.c_void_p)
        assert values.size == n0 * n1
        rc = self.L.orc_tables_set(self.T, table_id, n0, n1, x0, x1p, values.reshape(-1))
        assert rc == 0

    def evolve_batch(self, props, flags, time_end, n_threads: int = 1):
        n = props.shape[0]
        assert props.shape == (n, abi.NPROP) and props.dtype == np.float64
        status = np.zeros(n, dtype=np.int32)
        interrupt = np.zeros(n, dtype=np.int32)
        c = abi.glc_counters()
        te = np.ascontiguousarray(time_end, dtype=np.float64)
        rc = self.L.orc_evolve_batch(C.byref(self.params), self.T, n, props, flags, te, status,
                                     interrupt, C.byref(c), n_threads)
        assert rc == 0
        return status, interrupt, abi.counters_dict(c)

    def profiler_reset(self):
        self.L.orc_profiler_reset(C.byref(self.params))

    def profiler_read(self):
        pr = abi.glc_profile()
        self.L.orc_profiler_read(C.byref(pr))
        return abi.profile_dict(pr)

    def error_report(self, record, flag, time_step):
        """standardErrorHandler's table (orc_error_report): dict of [NY] arrays y, dydt, scale, tolerance, error, error_scaled,
        active, and the interrupt code of the evaluation at the node's time."""
        out = np.zeros(7 * abi.NY, dtype=np.float64)
        code = C.c_int(0)
        row = np.ascontiguousarray(record, dtype=np.float64).copy()
        self.L.orc_error_report.argtypes = [C.POINTER(abi.glc_params), C.c_void_p, np.ctypeslib.ndpointer(np.float64), C.c_int,
                                            C.c_double, np.ctypeslib.ndpointer(np.float64), C.POINTER(C.c_int)]
        self.L.orc_error_report.restype = C.c_int
        self.L.orc_error_report(C.byref(self.params), self.T, row, int(flag), float(time_step), out, C.byref(code))
        o = out.reshape(7, abi.NY)
        return {"y": o[0], "dydt": o[1], "scale": o[2], "tolerance": o[3], "error": o[4], "error_scaled": o[5],
                "active": o[6].astype(np.int32), "interrupt": code.value}

    def rhs(self, props_row, flag):
        dydt = np.zeros(abi.NY, dtype=np.float64)
        code = C.c_int(0)
        row = np.ascontiguousarray(props_row, dtype=np.float64).copy()
        self.L.orc_rhs_node(C.byref(self.params), self.T, row, int(flag), dydt, C.byref(code))
        return dydt, code.value, row

    def forest_evolve(self, forest, n_threads: int = 1):
        """Tree-level evolution (orc_tree.c): returns (records, flags, state, forest_counters, counters)."""
        n = forest["parent"].shape[0]
        rec = np.zeros((n, abi.NPROP))
        flags = np.zeros(n, dtype=np.int32)
        state = np.zeros(n, dtype=np.int32)
        fc, c = abi.glc_forest_counters(), abi.glc_counters()
        a = [np.ascontiguousarray(forest[k], dtype=np.float64) for k in ("mass", "time", "scale_radius", "angular_momentum")]
        rc = self.L.orc_forest_evolve(C.byref(self.params), self.T, n, np.ascontiguousarray(forest["parent"], dtype=np.int32),
                                      a[0], a[1], a[2], a[3], rec, flags, state, C.byref(fc), C.byref(c), n_threads)
        assert rc == 0, rc
        return rec, flags, state, abi.counters_dict(fc), abi.counters_dict(c)

