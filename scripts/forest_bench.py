"""Tree-level throughput: N Milky-Way-mass binary-split trees (root 1.52e12 Msun, resolution 1e9 Msun: the masses of
testSuite/parameters/benchmark_milkyWay.xml:30-42) through glc_forest_evolve, and a sample of them through the CPU checker's
tree walk (OpenMP over trees) on the host cores.
usage: python scripts/forest_bench.py N_TREES [CPU_SAMPLE_TREES] [KEY=VALUE env knobs ...]"""
import json
import os
import sys
import time

sys.path.insert(0, '.')
args = [a for a in sys.argv[1:] if '=' not in a]
for kv in sys.argv[1:]:
    if '=' in kv:
        k, v = kv.split('=', 1)
        os.environ[k] = v
n_trees = int(args[0])
cpu_trees = int(args[1]) if len(args) > 1 else 0
import numpy as np  # noqa: E402

import bench  # noqa: E402
from galacticus_b200 import abi, synthetic  # noqa: E402
from galacticus_b200.evolver import Evolver  # noqa: E402

p, _, _, _ = bench.workload(8, 219)
if os.environ.get("FOREST_KIND", "mw") == "volume":  # configs[3]: mass-function-sampled roots at the quickTest resolution
    f = synthetic.mass_function_forest(p, n_trees, 5.0e9, seed=219)
else:
    f = synthetic.binary_split_forest(p, n_trees, 1.52e12, 1.0e9, seed=219)
ev = Evolver(0)
synthetic.install(ev, p)
warm = synthetic.binary_split_forest(p, 4, 1.52e12, 1.0e10, seed=1)
ev.forest_evolve(warm)
t0 = time.perf_counter()
rec, flags, state, fc, c = ev.forest_evolve(f)
dt = time.perf_counter() - t0
out = {"trees": n_trees, "nodes": int(f["parent"].shape[0]), "seconds": dt, "trees_per_s": n_trees / dt,
       "node_ode_steps_per_s": c["steps_accepted"] / dt, "rhs_per_s": c["rhs_evaluations"] / dt, "forest_counters": fc,
       "counters": c, "knobs": [a for a in sys.argv[1:] if '=' in a]}
if cpu_trees:
    from oracle import orc

    orc.build()
    sub = synthetic.binary_split_forest(p, cpu_trees, 1.52e12, 1.0e9, seed=219)
    o = orc.Oracle(fast=True)
    synthetic.install(o, p)
    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    _, _, _, fco, co = o.forest_evolve(sub, n_threads=cores)
    dtc = time.perf_counter() - t0
    out["cpu"] = {"trees": cpu_trees, "cores": cores, "seconds": dtc, "trees_per_s": cpu_trees / dtc,
                  "node_ode_steps_per_s": co["steps_accepted"] / dtc}
print("FOREST", json.dumps(out), flush=True)
