"""GPU perf probe: evolve the bench workload under several execution options.  usage: perf_probe.py N [configs]"""
import os, sys, time, numpy as np
sys.path.insert(0, '.')
import bench
from galacticus_b200 import abi, synthetic
from galacticus_b200.evolver import Evolver
n = int(sys.argv[1])
configs = sys.argv[2:] or ["0:1", "256:1", "256:0"]
p, props, flags, tend = bench.workload(n, 219)
ev = Evolver(0); synthetic.install(ev, p)
ev.arena_upload(props, flags, tend); ev.arena_snapshot(n)
ref = None
for cfg in configs:
    budget, sort = [int(x) for x in cfg.split(":")]
    ev.set_option(abi.GLC_OPT_SLICE_BUDGET, budget); ev.set_option(abi.GLC_OPT_SORT_QUEUE, sort)
    for rep in range(2):
        ev.arena_restore(n)
        s0 = ev.slice_count()
        c, ms = ev.evolve_arena(n)
        print("cfg budget=%d sort=%d rep=%d: %.1f ms, slices %d, rhs/s %.3e steps/s %.3e nodes/s %.3e" % (
            budget, sort, rep, ms, ev.slice_count() - s0, c['rhs_evaluations'] / ms * 1e3, c['steps_accepted'] / ms * 1e3, n / ms * 1e3), flush=True)
    out = ev.arena_download(n)[0]
    if ref is None: ref = out
    else: print("   identical to first config:", np.array_equal(ref, out), flush=True)
