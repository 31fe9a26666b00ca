"""Critical path of a forest under the asynchronous schedule (test infrastructure: the host-executed kernel source of
tests/emu over a virtual clock with unlimited lanes).  usage: python scripts/forest_critical_path.py N_TREES [RESOLUTION] [OVERHEAD_EVALS]"""
import ctypes as C
import sys
import time

import numpy as np

sys.path.insert(0, '.')
import bench  # noqa: E402
from galacticus_b200 import abi, synthetic  # noqa: E402
from tests import emu  # noqa: E402

n_trees = int(sys.argv[1])
res = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0e9
overhead = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
emu.build()
p, _, _, _ = bench.workload(8, 219)
f = synthetic.binary_split_forest(p, n_trees, 1.52e12, res, seed=219)
E = emu.EmuEvolver(nslots=1, machine=False)
synthetic.install(E, p)
L = emu.lib()
_dp = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
L.emu_forest_critical_path.argtypes = [C.c_void_p, C.c_int64, _ip, _dp, _dp, _dp, _dp, _dp, _ip, _ip, C.c_double, _dp]
n = f["parent"].shape[0]
rec = np.zeros((n, abi.NPROP)); fl = np.zeros(n, np.int32); st = np.zeros(n, np.int32); out = np.zeros(3)
t0 = time.perf_counter()
rc = L.emu_forest_critical_path(E.h, n, f["parent"], f["mass"], f["time"], f["scale_radius"], f["angular_momentum"], rec, fl, st, overhead, out)
print("rc", rc, "trees", n_trees, "nodes", n, "critical path %.0f evaluations, total %.0f evaluations in %d evolves, mean in flight %.1f (%.1f per tree); %.1f s"
      % (out[0], out[1], out[2], out[1] / out[0], out[1] / out[0] / n_trees, time.perf_counter() - t0))
