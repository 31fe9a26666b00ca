"""Host-side handle on the C-ABI shared library ``libglcb200.so``.

This is a thin ctypes binding used by tests and ``bench.py``; the production host is the
Fortran shim shown in INTEGRATION.md (or the C++ mirror in ``csrc/host``), which binds the same
``extern "C"`` entry points.  There is NO CPU fallback: if the CUDA extension is missing or no
GPU is visible, construction fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import abi

LIB_PATH = os.environ.get("GLC_LIB_PATH") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libglcb200.so")
_LIB = None

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


class GlcError(RuntimeError):
    pass


def load_library() -> C.CDLL:
    """Load libglcb200.so (built in-tree by build.sh / __graft_entry__.build())."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise GlcError(
            f"{LIB_PATH} is missing: build it with ./build.sh (nvcc, sm_100a). "
            "galacticus_b200 has no CPU fallback."
        )
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.glc_abi_version.restype = C.c_int
    L.glc_last_error.restype = C.c_char_p
    L.glc_last_error.argtypes = [vp]
    L.glc_evolver_create.argtypes = [C.POINTER(vp), C.c_int32]
    L.glc_evolver_destroy.argtypes = [vp]
    L.glc_evolver_set_params.argtypes = [vp, C.POINTER(abi.glc_params)]
    L.glc_params_default.argtypes = [C.POINTER(abi.glc_params), C.c_int32]
    L.glc_evolver_set_table.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, _dp, vp, _dp]
    L.glc_evolve_batch.argtypes = [vp, C.c_int64, _dp, _ip, _dp, _ip, _ip, C.POINTER(abi.glc_counters)]
    L.glc_arena_reserve.argtypes = [vp, C.c_int64]
    L.glc_arena_upload.argtypes = [vp, C.c_int64, _dp, _ip, _dp]
    L.glc_arena_download.argtypes = [vp, C.c_int64, vp, vp, vp, vp]
    L.glc_evolve_arena.argtypes = [vp, C.c_int64, C.POINTER(abi.glc_counters)]
    L.glc_last_kernel_ms.restype = C.c_float
    L.glc_last_kernel_ms.argtypes = [vp]
    if hasattr(L, "glc_profiler_read"):
        L.glc_profiler_read.argtypes = [vp, C.POINTER(abi.glc_profile)]
        L.glc_profiler_reset.argtypes = [vp]
    if hasattr(L, "glc_last_phase_stats"):  # (absent from older builds kept for regression experiments)
        L.glc_last_phase_stats.argtypes = [vp, _dp]
    L.glc_arena_device_props.restype = vp
    L.glc_arena_device_props.argtypes = [vp]
    L.glc_evolver_stream.restype = vp
    L.glc_evolver_stream.argtypes = [vp]
    L.glc_rhs_batch.argtypes = [vp, C.c_int64, _dp, _ip, _dp, _ip]
    L.glc_forest_evolve.argtypes = [vp, C.c_int64, _ip, _dp, _dp, _dp, _dp, _dp, _ip, _ip,
                                    C.POINTER(abi.glc_forest_counters), C.POINTER(abi.glc_counters)]
    L.glc_arena_snapshot.argtypes = [vp, C.c_int64]
    L.glc_arena_restore.argtypes = [vp, C.c_int64]
    L.glc_arena_capacity.restype = C.c_int64
    L.glc_arena_capacity.argtypes = [vp]
    L.glc_kernel_launch_count.restype = C.c_int64
    L.glc_kernel_launch_count.argtypes = [vp]
    L.glc_measure_fp64_peak_tflops.restype = C.c_double
    L.glc_measure_fp64_peak_tflops.argtypes = [vp]
    L.glc_evolver_set_option.argtypes = [vp, C.c_int32, C.c_int64]
    L.glc_slice_count.restype = C.c_int64
    L.glc_slice_count.argtypes = [vp]
    i64p = C.POINTER(C.c_int64)
    L.glc_stream_begin.argtypes = [vp, C.c_int64]
    L.glc_stream_submit.argtypes = [vp, C.c_int64, _dp, _ip, _dp, i64p]
    L.glc_stream_run.argtypes = [vp, C.c_int32, i64p, C.POINTER(abi.glc_counters)]
    L.glc_stream_collect.argtypes = [vp, C.c_int64, np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS"), _dp, _ip, _ip,
                                     _ip, i64p]
    L.glc_stream_finish.argtypes = [vp, C.POINTER(abi.glc_counters)]
    L.glc_stream_end.argtypes = [vp]
    L.glc_histogram_accumulate.argtypes = [vp, C.c_int64, C.c_int32, C.c_double, C.c_double, C.c_int32, vp]
    if L.glc_abi_version() != abi.GLC_ABI_VERSION:
        raise GlcError("libglcb200.so ABI version does not match include/glc_b200.h")
    _LIB = L
    return L


def params_default(model: int) -> abi.glc_params:
    p = abi.glc_params()
    rc = load_library().glc_params_default(C.byref(p), model)
    if rc != 0:
        raise GlcError(f"glc_params_default failed ({rc})")
    return p


class Evolver:
    """Batched ``mergerTreeNodeEvolver`` on one GPU (mirrors
    source/merger_trees/node_evolver/_class.F90:34-82, batched)."""

    def __init__(self, device: int = 0):
        self.L = load_library()
        self.h = C.c_void_p()
        rc = self.L.glc_evolver_create(C.byref(self.h), device)
        if rc != 0:
            raise GlcError(f"glc_evolver_create(device={device}) failed ({rc}): no usable CUDA device? "
                           "galacticus_b200 has no CPU fallback.")

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.L.glc_evolver_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str):
        if rc > 0:  # completed with a warning (GLC_WARN_*): the outputs are valid, the caller inspects the counters
            import warnings

            msg = self.L.glc_last_error(self.h)
            warnings.warn(f"{what}: {msg.decode() if msg else rc}")
        elif rc != 0:
            msg = self.L.glc_last_error(self.h)
            raise GlcError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")

    def set_params(self, p: abi.glc_params) -> None:
        self.params = p
        self._check(self.L.glc_evolver_set_params(self.h, C.byref(p)), "glc_evolver_set_params")

    def set_table(self, table_id: int, x0, x1, values) -> None:
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        values = np.ascontiguousarray(values, dtype=np.float64).reshape(-1)
        n0 = x0.size
        if x1 is None:
            n1 = values.size // n0
            x1p = None
        else:
            x1 = np.ascontiguousarray(x1, dtype=np.float64)
            n1 = x1.size
            x1p = x1.ctypes.data_as(C.c_void_p)
        assert values.size == n0 * n1
        self._check(self.L.glc_evolver_set_table(self.h, table_id, n0, n1, x0, x1p, values), "glc_evolver_set_table")

    def evolve_batch(self, props, flags, time_end):
        """In-place on host arrays; returns (status, interrupt, counters)."""
        n = props.shape[0]
        assert props.shape == (n, abi.NPROP) and props.dtype == np.float64 and props.flags.c_contiguous
        status = np.zeros(n, dtype=np.int32)
        interrupt = np.zeros(n, dtype=np.int32)
        c = abi.glc_counters()
        te = np.ascontiguousarray(time_end, dtype=np.float64)
        self._check(self.L.glc_evolve_batch(self.h, n, props, flags, te, status, interrupt, C.byref(c)), "glc_evolve_batch")
        return status, interrupt, abi.counters_dict(c)

    # ---- device-resident path
    def arena_upload(self, props, flags, time_end):
        n = props.shape[0]
        te = np.ascontiguousarray(time_end, dtype=np.float64)
        self._check(self.L.glc_arena_upload(self.h, n, props, flags, te), "glc_arena_upload")

    def evolve_arena(self, n: int):
        c = abi.glc_counters()
        self._check(self.L.glc_evolve_arena(self.h, n, C.byref(c)), "glc_evolve_arena")
        return abi.counters_dict(c), float(self.L.glc_last_kernel_ms(self.h))

    def profiler_read(self):
        """mergerTreeEvolveProfilerSimple accumulators (profileOdeEvolver must be set) as a dict of numpy arrays."""
        pr = abi.glc_profile()
        self._check(self.L.glc_profiler_read(self.h, C.byref(pr)), "glc_profiler_read")
        return abi.profile_dict(pr)

    def error_report_node(self, record, flag, time_step):
        """standardErrorHandler's "ODE system parameters" table for one node (glc_error_report_node) as a dict of [NY] arrays."""
        rep = abi.glc_error_report()
        row = np.ascontiguousarray(record, dtype=np.float64)
        self.L.glc_error_report_node.argtypes = [C.c_void_p, np.ctypeslib.ndpointer(np.float64), C.c_int32, C.c_double,
                                                 C.POINTER(abi.glc_error_report)]
        self.L.glc_error_report_node.restype = C.c_int
        self._check(self.L.glc_error_report_node(self.h, row, int(flag), float(time_step), C.byref(rep)), "glc_error_report_node")
        f = lambda a: np.array(a[:], dtype=np.float64)  # noqa: E731
        return {"y": f(rep.y), "dydt": f(rep.dydt), "scale": f(rep.scale), "tolerance": f(rep.tolerance), "error": f(rep.error),
                "error_scaled": f(rep.error_scaled), "active": np.array(rep.active[:], dtype=np.int32), "interrupt": int(rep.interrupt),
                "time": float(rep.time), "time_step": float(rep.time_step)}

    def profiler_reset(self) -> None:
        self._check(self.L.glc_profiler_reset(self.h), "glc_profiler_reset")

    def last_phase_stats(self):
        """{machine_ms, drain_ms, machine_rhs, drain_rhs, machine_steps, drain_steps} of the last machine batch."""
        out = np.zeros(8)
        self._check(self.L.glc_last_phase_stats(self.h, out), "glc_last_phase_stats")
        return dict(zip(("machine_ms", "drain_ms", "machine_rhs", "drain_rhs", "machine_steps", "drain_steps", "machine_nodes",
                         "drain_nodes"), out.tolist()))

    def arena_download(self, n: int):
        props = np.zeros((n, abi.NPROP), dtype=np.float64)
        flags = np.zeros(n, dtype=np.int32)
        status = np.zeros(n, dtype=np.int32)
        interrupt = np.zeros(n, dtype=np.int32)
        vp = C.c_void_p
        self._check(self.L.glc_arena_download(self.h, n, props.ctypes.data_as(vp), flags.ctypes.data_as(vp),
                                              status.ctypes.data_as(vp), interrupt.ctypes.data_as(vp)),
                    "glc_arena_download")
        return props, flags, status, interrupt

    def arena_snapshot(self, n: int):
        self._check(self.L.glc_arena_snapshot(self.h, n), "glc_arena_snapshot")

    def arena_restore(self, n: int):
        self._check(self.L.glc_arena_restore(self.h, n), "glc_arena_restore")

    def set_option(self, option: int, value: int) -> None:
        self._check(self.L.glc_evolver_set_option(self.h, option, value), "glc_evolver_set_option")

    def slice_count(self) -> int:
        return int(self.L.glc_slice_count(self.h))

    def kernel_launch_count(self) -> int:
        return int(self.L.glc_kernel_launch_count(self.h))

    def fp64_peak_tflops(self) -> float:
        return float(self.L.glc_measure_fp64_peak_tflops(self.h))

    def device_props_ptr(self) -> int:
        return int(self.L.glc_arena_device_props(self.h) or 0)

    # ---- streaming session (glc_stream_*): submit / run time slices / collect finished nodes
    def stream_begin(self, capacity: int) -> None:
        self._check(self.L.glc_stream_begin(self.h, capacity), "glc_stream_begin")

    def stream_submit(self, props, flags, time_end) -> int:
        n = props.shape[0]
        assert props.shape == (n, abi.NPROP) and props.dtype == np.float64 and props.flags.c_contiguous
        first = C.c_int64(-1)
        te = np.ascontiguousarray(time_end, dtype=np.float64)
        self._check(self.L.glc_stream_submit(self.h, n, props, np.ascontiguousarray(flags, dtype=np.int32), te, C.byref(first)),
                    "glc_stream_submit")
        return int(first.value)

    def stream_run(self, pops_per_warp: int = 0):
        done = C.c_int64(0)
        c = abi.glc_counters()
        self._check(self.L.glc_stream_run(self.h, pops_per_warp, C.byref(done), C.byref(c)), "glc_stream_run")
        return int(done.value), abi.counters_dict(c)

    def stream_collect(self, max_nodes: int):
        tickets = np.zeros(max_nodes, dtype=np.int64)
        props = np.zeros((max_nodes, abi.NPROP), dtype=np.float64)
        flags = np.zeros(max_nodes, dtype=np.int32)
        status = np.zeros(max_nodes, dtype=np.int32)
        interrupt = np.zeros(max_nodes, dtype=np.int32)
        m = C.c_int64(0)
        self._check(self.L.glc_stream_collect(self.h, max_nodes, tickets, props, flags, status, interrupt, C.byref(m)),
                    "glc_stream_collect")
        k = int(m.value)
        return tickets[:k], props[:k], flags[:k], status[:k], interrupt[:k]

    def stream_finish(self):
        c = abi.glc_counters()
        self._check(self.L.glc_stream_finish(self.h, C.byref(c)), "glc_stream_finish")
        return abi.counters_dict(c)

    def stream_end(self) -> None:
        self._check(self.L.glc_stream_end(self.h), "glc_stream_end")

    def rhs_batch(self, props, flags):
        n = props.shape[0]
        dydt = np.zeros((n, abi.NY), dtype=np.float64)
        interrupt = np.zeros(n, dtype=np.int32)
        p = props.copy()
        self._check(self.L.glc_rhs_batch(self.h, n, p, flags, dydt, interrupt), "glc_rhs_batch")
        return dydt, interrupt, p

    def forest_evolve(self, forest):
        """Tree-level evolution of a set of forests (glc_forest_evolve: the batching tree evolver of INTEGRATION.md section 3).
        forest: dict of flat arrays parent / mass / time / scale_radius / angular_momentum (synthetic.binary_split_forest).
        Returns (records, flags, state, forest_counters, counters)."""
        n = forest["parent"].shape[0]
        rec = np.zeros((n, abi.NPROP))
        flags = np.zeros(n, dtype=np.int32)
        state = np.zeros(n, dtype=np.int32)
        fc, c = abi.glc_forest_counters(), abi.glc_counters()
        a = [np.ascontiguousarray(forest[k], dtype=np.float64) for k in ("mass", "time", "scale_radius", "angular_momentum")]
        self._check(self.L.glc_forest_evolve(self.h, n, np.ascontiguousarray(forest["parent"], dtype=np.int32), a[0], a[1], a[2],
                                             a[3], rec, flags, state, C.byref(fc), C.byref(c)), "glc_forest_evolve")
        return rec, flags, state, abi.counters_dict(fc), abi.counters_dict(c)
