import sys, numpy as np
sys.path.insert(0, '.')
from galacticus_b200 import abi, synthetic
from galacticus_b200.evolver import Evolver
from oracle import orc
from tests import cases
P = abi.P
names = {v: k for k, v in P.items()}
p = cases.standard_params()
ev = Evolver(0); synthetic.install(ev, p)
o = orc.Oracle(); synthetic.install(o, p)
n = 4000
props, flags, tend = synthetic.standard_nodes(p, n, seed=5)
pg2, fg2 = props.copy(), flags.copy(); po2, fo2 = props.copy(), flags.copy()
sg, ig2, cg = ev.evolve_batch(pg2, fg2, tend)
so, io2, co = o.evolve_batch(po2, fo2, tend, n_threads=8)
ne = (pg2 != po2).any(axis=1)
idx = np.where(ne)[0]
print("differing nodes", idx.size, "flags-changed among them", (fg2[idx] != flags[idx]).sum(), "of all changed", (fg2 != flags).sum())
for i in idx[:6]:
    q = props[i:i+1].copy(); f = flags[i:i+1].copy(); qo = q.copy(); fo = f.copy()
    s1, i1, c1 = ev.evolve_batch(q, f, tend[i:i+1]); s2, i2, c2 = o.evolve_batch(qo, fo, tend[i:i+1])
    d = np.where(q[0] != qo[0])[0]
    print(i, bin(flags[i]), "->", bin(f[0]), "t0 %.4f tend %.4f tstep %.4g" % (props[i, P['TIME']], tend[i], props[i, P['TIME_STEP']]), c1, c2)
    rel = np.abs(q[0] - qo[0]) / (np.abs(qo[0]) + 1e-300)
    print("    ", [(names[j], "%.2e" % rel[j]) for j in d[:8]])
