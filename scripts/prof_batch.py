"""Profiling harness: evolve a batch of node records whose CPU cost is below a cut (so that ncu's ~40 replays
of the kernel stay short).  usage: prof_batch.py N MAX_RHS [repeat]"""
import sys, time, numpy as np
sys.path.insert(0, '.')
from galacticus_b200 import abi, synthetic
from galacticus_b200.evolver import Evolver
from oracle import orc
from tests import cases
n, max_rhs = int(sys.argv[1]), int(sys.argv[2])
rep = int(sys.argv[3]) if len(sys.argv) > 3 else 1
p = cases.standard_params()
props, flags, tend = synthetic.standard_nodes(p, 3 * n, seed=219)
o = orc.Oracle(fast=True); synthetic.install(o, p)
keep = []
for i in range(props.shape[0]):
    q = props[i:i+1].copy(); f = flags[i:i+1].copy()
    _, _, c = o.evolve_batch(q, f, tend[i:i+1])
    if c['rhs_evaluations'] <= max_rhs: keep.append(i)
    if len(keep) == n: break
keep = np.array(keep)
props, flags, tend = np.ascontiguousarray(props[keep]), np.ascontiguousarray(flags[keep]), np.ascontiguousarray(tend[keep])
ev = Evolver(0); synthetic.install(ev, p)
ev.arena_upload(props, flags, tend); ev.arena_snapshot(len(keep))
for r in range(rep):
    ev.arena_restore(len(keep))
    c, ms = ev.evolve_arena(len(keep))
    print("n", len(keep), "kernel ms", ms, c, "rhs/s %.3e" % (c['rhs_evaluations'] / ms * 1e3))
