import sys, numpy as np
sys.path.insert(0, '.')
from galacticus_b200 import abi, synthetic
from galacticus_b200.evolver import Evolver
from oracle import orc
from tests import cases
P = abi.P
p = cases.standard_params()
ev = Evolver(0); synthetic.install(ev, p)
o = orc.Oracle(); synthetic.install(o, p)
n = 4000
props, flags, tend = synthetic.standard_nodes(p, n, seed=5)
dg, ig, pg = ev.rhs_batch(props, flags)
names = {v: k for k, v in P.items()}
do = np.zeros_like(dg); po = props.copy(); io = np.zeros_like(ig)
for i in range(n):
    do[i], io[i], po[i] = o.rhs(props[i], flags[i])
print("interrupt mismatches", (ig != io).sum())
neq = (dg != do)
print("non-bit-identical dydt entries:", neq.sum(), "of", dg.size, "nodes affected", neq.any(axis=1).sum())
cols = neq.sum(axis=0)
print({names[j]: int(cols[j]) for j in range(abi.NY) if cols[j]})
for k in ("DISK_RADIUS", "DISK_VELOCITY", "SPH_RADIUS", "SPH_VELOCITY", "BASIC_MASS"):
    print(k, "non-identical:", int((pg[:, P[k]] != po[:, P[k]]).sum()))
rel = np.abs(dg - do) / (np.maximum(np.abs(dg), np.abs(do)) + 1e-300)
print("max rel err", rel.max())
bad = np.argwhere(neq)
for (i, j) in bad[:12]:
    print(i, names[j], repr(dg[i, j]), repr(do[i, j]), bin(flags[i]))
# full evolution
pg2, fg2 = props.copy(), flags.copy(); po2, fo2 = props.copy(), flags.copy()
sg, ig2, cg = ev.evolve_batch(pg2, fg2, tend)
so, io2, co = o.evolve_batch(po2, fo2, tend, n_threads=8)
print("evolve: counters equal", cg == co, cg, co)
print("status eq", (sg == so).all(), "int eq", (ig2 == io2).all(), "flags eq", (fg2 == fo2).all())
ne = pg2 != po2
print("evolve non-identical entries", ne.sum(), "nodes", ne.any(axis=1).sum())
rel = np.abs(pg2 - po2) / (np.maximum(np.abs(pg2), np.abs(po2)) + 1e-300)
print("evolve max rel", rel.max(), {names[j]: int(ne[:, j].sum()) for j in range(abi.NPROP) if ne[:, j].any()})
