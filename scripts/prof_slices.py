"""Profiling harness: evolve the first GLC_MAX_SLICES time slices of the bench workload (all lanes busy).
usage: GLC_SLICE_BUDGET=64 GLC_MAX_SLICES=2 GLC_SLICE_LOG=1 prof_slices.py N"""
import sys, numpy as np
sys.path.insert(0, '.')
import bench
from galacticus_b200 import abi, synthetic
from galacticus_b200.evolver import Evolver
n = int(sys.argv[1])
p, props, flags, tend = bench.workload(n, 219)
ev = Evolver(0); synthetic.install(ev, p)
ev.arena_upload(props, flags, tend)
c, ms = ev.evolve_arena(n)
print("n", n, "ms", ms, c, "rhs/s %.3e" % (c['rhs_evaluations'] / ms * 1e3))
