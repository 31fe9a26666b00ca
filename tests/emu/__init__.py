"""TEST INFRASTRUCTURE ONLY: ctypes front end to tests/_build/libglcemu.so, the g++ build of the CUDA
kernels' per-lane logic (see glc_emu.cpp).  Never imported by the product package."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from galacticus_b200 import abi

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SO = os.path.join(os.path.dirname(HERE), "_build", "libglcemu.so")
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_LIB = None


def build(force: bool = False) -> None:
    src = os.path.join(HERE, "glc_emu.cpp")
    csrc = os.path.join(ROOT, "galacticus_b200", "csrc")
    deps = [src, abi.HEADER] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cuh", ".h"))]
    deps += [os.path.join(csrc, "host", f) for f in os.listdir(os.path.join(csrc, "host"))]
    if not force and os.path.exists(SO) and all(os.path.getmtime(d) <= os.path.getmtime(SO) for d in deps):
        return
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    subprocess.run(["g++", "-std=c++17", "-O2", "-mfma", "-ffp-contract=off", "-fPIC", "-shared", "-Wall", "-Wno-unused",
                    "-Wno-unknown-pragmas", "-o", SO, src], check=True)


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        build()
        L = C.CDLL(SO)
        L.emu_create.restype = C.c_void_p
        L.emu_destroy.argtypes = [C.c_void_p]
        L.emu_set_params.argtypes = [C.c_void_p, C.POINTER(abi.glc_params)]
        L.emu_set_table.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, _dp, C.c_void_p, _dp]
        L.emu_evolve_batch.argtypes = [C.c_void_p, C.c_int64, _dp, _ip, _dp, _ip, _ip, C.POINTER(abi.glc_counters),
                                       C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int64)]
        L.emu_forest_evolve.argtypes = [C.c_void_p, C.c_int64, _ip, _dp, _dp, _dp, _dp, _dp, _ip, _ip,
                                        C.POINTER(abi.glc_forest_counters), C.POINTER(abi.glc_counters), C.c_int, C.c_int, C.c_int,
                                        C.c_int]
        L.emu_forest_evolve_async.argtypes = [C.c_void_p, C.c_int64, _ip, _dp, _dp, _dp, _dp, _dp, _ip, _ip,
                                              C.POINTER(abi.glc_forest_counters), C.POINTER(abi.glc_counters), C.c_int, C.c_int, C.c_int,
                                              C.c_int, C.c_int]
        L.emu_profiler_read.argtypes = [C.c_void_p, C.POINTER(abi.glc_profile)]
        _LIB = L
    return _LIB


class EmuEvolver:
    """Same calling convention as galacticus_b200.Evolver.evolve_batch, executed lane by lane on the host."""

    def __init__(self, nslots: int = 64, budget: int = 0, sort: bool = True, machine=True):
        """machine: False = lane kernel logic, True/1 = micro-task machine, 2 = machine + drain hand-over."""
        self.L = lib()
        self.h = C.c_void_p(self.L.emu_create())
        self.nslots, self.budget, self.sort, self.machine = nslots, budget, sort, machine
        self.slices = 0
        self._keep = []

    def __del__(self):
        try:
            self.L.emu_destroy(self.h)
        except Exception:
            pass

    def set_params(self, p: abi.glc_params) -> None:
        self.L.emu_set_params(self.h, C.byref(p))

    def set_table(self, table_id: int, x0, x1, values) -> None:
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        values = np.ascontiguousarray(values, dtype=np.float64).reshape(-1)
        n0 = x0.size
        if x1 is None:
            n1, x1p = values.size // n0, None
        else:
            x1 = np.ascontiguousarray(x1, dtype=np.float64)
            n1, x1p = x1.size, x1.ctypes.data_as(C.c_void_p)
        assert self.L.emu_set_table(self.h, table_id, n0, n1, x0, x1p, values) == 0

    def evolve_batch(self, props, flags, time_end):
        n = props.shape[0]
        status = np.zeros(n, dtype=np.int32)
        interrupt = np.zeros(n, dtype=np.int32)
        c = abi.glc_counters()
        s = C.c_int64(0)
        te = np.ascontiguousarray(time_end, dtype=np.float64)
        rc = self.L.emu_evolve_batch(self.h, n, props, flags, te, status, interrupt, C.byref(c), self.nslots,
                                     self.budget, int(self.sort), int(self.machine), C.byref(s))
        assert rc == 0
        self.slices = s.value
        return status, interrupt, abi.counters_dict(c)

    def stream_session(self, props, flags, time_end, chunk, pattern, lane_budget=6, machine_budget=5):
        """A streaming session with adaptive ticks on the host (emu_stream_session): nodes submitted in chunks, after each chunk
        the ticks of `pattern` ('M' = machine slice, 'L' = lane pass with refill), finished by the machine.  Returns
        (rc, status, interrupt, counters); rc 0 = every node written back exactly once."""
        n = props.shape[0]
        status = np.zeros(n, dtype=np.int32)
        interrupt = np.zeros(n, dtype=np.int32)
        c = abi.glc_counters()
        te = np.ascontiguousarray(time_end, dtype=np.float64)
        self.L.emu_stream_session.argtypes = [C.c_void_p, C.c_int64, _dp, _ip, _dp, _ip, _ip, C.POINTER(abi.glc_counters),
                                              C.c_int, C.c_int, C.c_char_p, C.c_int, C.c_int]
        self.L.emu_stream_session.restype = C.c_int
        rc = self.L.emu_stream_session(self.h, n, props, flags, te, status, interrupt, C.byref(c), self.nslots, int(chunk),
                                       pattern.encode(), int(lane_budget), int(machine_budget))
        return rc, status, interrupt, abi.counters_dict(c)

    def profiler_read(self):
        pr = abi.glc_profile()
        assert self.L.emu_profiler_read(self.h, C.byref(pr)) == 0
        return abi.profile_dict(pr)

    def forest_evolve(self, forest, asynchronous=False, straggle=0):
        """The product's host scheduler (csrc/host/glc_forest.hpp) over the host-executed kernel source: the bulk-synchronous
        rounds, or (asynchronous) the per-group schedule with finished nodes reported up to `straggle` polls late."""
        n = forest["parent"].shape[0]
        rec = np.zeros((n, abi.NPROP))
        flags = np.zeros(n, dtype=np.int32)
        state = np.zeros(n, dtype=np.int32)
        fc, c = abi.glc_forest_counters(), abi.glc_counters()
        a = [np.ascontiguousarray(forest[k], dtype=np.float64) for k in ("mass", "time", "scale_radius", "angular_momentum")]
        if asynchronous:
            rc = self.L.emu_forest_evolve_async(self.h, n, np.ascontiguousarray(forest["parent"], dtype=np.int32), a[0], a[1], a[2],
                                                a[3], rec, flags, state, C.byref(fc), C.byref(c), self.nslots, self.budget,
                                                int(self.sort), int(self.machine), int(straggle))
        else:
            rc = self.L.emu_forest_evolve(self.h, n, np.ascontiguousarray(forest["parent"], dtype=np.int32), a[0], a[1], a[2], a[3],
                                          rec, flags, state, C.byref(fc), C.byref(c), self.nslots, self.budget, int(self.sort),
                                          int(self.machine))
        assert rc == 0, rc
        return rec, flags, state, abi.counters_dict(fc), abi.counters_dict(c)

